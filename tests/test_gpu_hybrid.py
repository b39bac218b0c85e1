"""GPU: hybrid EKF-SLAM feature rows at the stage level through the C ABI against the oracle
restatement (oracle/hybrid.py, itself pinned by central differences in tests/test_oracle_hybrid_cpu.py)."""
import numpy as np
import pytest

from oracle import hybrid as hy
from oracle import mathutils as mu
from orcvio_b200 import api

pytestmark = pytest.mark.gpu


def _window(seed, N):
    rng = np.random.default_rng(seed)
    R_b2c, t_c_b = mu.so3_exp(rng.normal(0, 0.8, 3)), rng.normal(0, 0.1, 3)
    clone_R = [mu.so3_exp(rng.normal(0, 0.08, 3)) for _ in range(N)]
    clone_p = [np.array([0.25 * i, 0.02 * i, 0.0]) + rng.normal(0, 0.03, 3) for i in range(N)]
    return rng, R_b2c, t_c_b, clone_R, clone_p


def _features(rng, R_b2c, t_c_b, clone_R, clone_p, F, ragged=True):
    N = len(clone_R)
    anchor = rng.integers(0, N, F).astype(np.int32)
    rho = 1.0 / rng.uniform(3.0, 20.0, F)
    f_an = np.stack([rng.uniform(-0.4, 0.4, F), rng.uniform(-0.3, 0.3, F)], axis=1)
    pos = np.array([hy.feature_position_from_anchor(clone_R[a], clone_p[a], R_b2c, t_c_b, [fx, fy, 1.0], r)
                    for a, r, (fx, fy) in zip(anchor, rho, f_an)])
    feat_off, obs_clone, obs_z = [0], [], []
    for f in range(F):
        m = int(rng.integers(1, N + 1)) if ragged else N
        cl = np.sort(rng.choice(N, m, replace=False))
        if f % 3 == 0 and anchor[f] not in cl:
            cl = np.sort(np.append(cl, anchor[f]))            # include the anchor's own observation (zeroed rows)
        for c in cl:
            p_ck = R_b2c @ clone_R[c].T @ (pos[f] - (clone_p[c] + clone_R[c] @ t_c_b))
            obs_clone.append(int(c))
            obs_z.append(p_ck[:2] / p_ck[2] + rng.normal(0, 0.004, 2))
        feat_off.append(len(obs_clone))
    return anchor, rho, f_an, pos, np.array(feat_off, np.int32), np.array(obs_clone, np.int32), np.array(obs_z)


@pytest.mark.parametrize("seed,N,F", [(0, 6, 9), (1, 20, 30), (2, 30, 64)])
def test_ekf_measurement_jacobians(seed, N, F):
    rng, R_b2c, t_c_b, clone_R, clone_p = _window(seed, N)
    anchor, rho, f_an, pos, feat_off, obs_clone, obs_z = _features(rng, R_b2c, t_c_b, clone_R, clone_p, F)
    out = api.ekf_measurement_jacobians(clone_R, clone_p, R_b2c, t_c_b, anchor, rho, f_an, pos, feat_off, obs_clone, obs_z)
    n_zero = 0
    for f in range(F):
        a = int(anchor[f])
        for o in range(feat_off[f], feat_off[f + 1]):
            c = int(obs_clone[o])
            ref = hy.measurement_jacobian_ekf_1didp(clone_R[c], clone_p[c], clone_R[a], clone_p[a], R_b2c, t_c_b,
                                                    [f_an[f, 0], f_an[f, 1], 1.0], rho[f], pos[f], obs_z[o],
                                                    same_state=(c == a))
            n_zero += c == a
            for got, want in zip((out["H_f"][o].reshape(2, 1), out["H_a"][o], out["H_x"][o], out["H_e"][o], out["r"][o]), ref):
                scale = max(np.abs(want).max(), 1e-300)
                assert np.abs(got - want).max() <= 1e-12 * max(scale, 1.0), (f, o)
    assert n_zero > 0                                           # the zeroed anchor-frame case was exercised


@pytest.mark.parametrize("seed,N,F", [(3, 8, 5), (4, 20, 30)])
def test_ekf_feature_rows_and_gate(seed, N, F):
    rng, R_b2c, t_c_b, clone_R, clone_p = _window(seed, N)
    anchor, rho, f_an, pos, _, _, _ = _features(rng, R_b2c, t_c_b, clone_R, clone_p, F)
    anchor[0] = N - 1                                           # anchored in the observing clone itself: zero rows
    pos[0] = hy.feature_position_from_anchor(clone_R[N - 1], clone_p[N - 1], R_b2c, t_c_b, [f_an[0, 0], f_an[0, 1], 1.0], rho[0])
    k = N - 1
    z = np.zeros((F, 2))
    for f in range(F):
        p_ck = R_b2c @ clone_R[k].T @ (pos[f] - (clone_p[k] + clone_R[k] @ t_c_b))
        z[f] = p_ck[:2] / p_ck[2] + rng.normal(0, 0.01 if f % 4 else 3.0, 2)   # every fourth one is a gross outlier
    D = 22 + 6 * N + F
    A = rng.normal(0, 0.03, (D, D))
    P = A @ A.T * 0.1 + 1e-5 * np.eye(D)
    P[15:22, :] = 0.0
    P[:, 15:22] = 0.0
    sigma2 = 6.4e-5
    out = api.ekf_feature_rows(clone_R, clone_p, R_b2c, t_c_b, anchor, rho, f_an, pos, z, P, sigma2, 0.95)
    chi2 = mu.chi2_table(0.95)[2]
    n_pass = 0
    for f in range(F):
        H, r = hy.feature_jacobian_ekf(clone_R, clone_p, R_b2c, t_c_b, k, int(anchor[f]), f, F,
                                       [f_an[f, 0], f_an[f, 1], 1.0], rho[f], pos[f], z[f])
        g, ok = hy.gate_ekf_row(H, r, P, sigma2, chi2)
        assert np.abs(out["H"][2 * f:2 * f + 2] - H).max() <= 1e-12 * max(np.abs(H).max(), 1.0)
        assert np.abs(out["r"][2 * f:2 * f + 2] - r).max() <= 1e-12
        assert abs(out["gamma"][f] - g) <= 1e-9 * max(abs(g), 1e-300)
        if abs(g - chi2) > 1e-9 * chi2:
            assert bool(out["pass"][f]) == ok
        n_pass += ok
    assert 0 < n_pass < F
    assert np.all(out["H"][0:2] == 0) and out["gamma"][0] == 0.0


@pytest.mark.parametrize("seed,N,E", [(5, 6, 3), (6, 20, 30)])
def test_anchor_change_covariance(seed, N, E):
    """H4: updateFeatureCov_1didp and rmLostFeaturesCov against the oracle."""
    rng, R_b2c, t_c_b, clone_R, clone_p = _window(seed, N)
    D = 22 + 6 * N + E
    A = rng.normal(0, 0.04, (D, D))
    P = A @ A.T + 1e-5 * np.eye(D)
    P[15:22, :] = 0.0
    P[:, 15:22] = 0.0
    for fidx, old, new in [(0, 0, N - 1), (E - 1, N - 2, 1), (E // 2, 2, 2 if N < 4 else 3)]:
        f_an, rho = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.2, 0.2), 1.0]), 1.0 / rng.uniform(4.0, 15.0)
        p_w = hy.feature_position_from_anchor(clone_R[old], clone_p[old], R_b2c, t_c_b, f_an, rho)
        p_c = R_b2c @ clone_R[new].T @ (p_w - (clone_p[new] + clone_R[new] @ t_c_b))
        rho_new = 1.0 / p_c[2]
        ref, J_ref = hy.update_feature_cov_1didp(P, N, fidx, old, new, clone_R, clone_p, R_b2c, t_c_b, p_w, rho_new)
        got, J = api.ekf_update_feature_cov(P, clone_R, clone_p, R_b2c, t_c_b, fidx, old, new, p_w, rho_new)
        assert np.abs(J - J_ref[0]).max() <= 1e-12 * max(np.abs(J_ref).max(), 1.0)
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
        assert np.abs(got - got.T).max() == 0.0
        P = ref                                                  # chain the anchor changes
    out = api.ekf_remove_feature_cov(P, N, E // 2)
    c = 22 + 6 * N + E // 2
    np.testing.assert_array_equal(out, np.delete(np.delete(P, c, axis=0), c, axis=1))


@pytest.mark.parametrize("seed,N,F", [(7, 7, 3), (8, 20, 24)])
def test_new_feature_rows_and_delayed_initialisation(seed, N, F):
    """H2 (featureJacobian_ekf_new) + H3 (sparsification) + the new-state part of measurementUpdate_hybrid.  The
    nullspace basis is not unique: the feature-free rows are compared through their Gram invariants,
    the initialisation through HH = H_2^-1 H_1, H_2^-1 r_1 and the augmented covariance, which are basis independent."""
    rng, R_b2c, t_c_b, clone_R, clone_p = _window(seed, N)
    anchor, rho, f_an, pos, feat_off, obs_clone, obs_z = _features(rng, R_b2c, t_c_b, clone_R, clone_p, F)
    # every feature needs at least two observations that are not the anchor's own
    keep = [f for f in range(F) if (obs_clone[feat_off[f]:feat_off[f + 1]] != anchor[f]).sum() >= 2]
    assert len(keep) >= 2
    fo, oc, oz = [0], [], []
    for f in keep:
        oc.extend(obs_clone[feat_off[f]:feat_off[f + 1]])
        oz.extend(obs_z[feat_off[f]:feat_off[f + 1]])
        fo.append(len(oc))
    anchor, rho, f_an, pos = anchor[keep], rho[keep], f_an[keep], pos[keep]
    feat_off, obs_clone, obs_z = np.array(fo, np.int32), np.array(oc, np.int32), np.array(oz)
    Fk = len(keep)
    D = 22 + 6 * N
    out = api.ekf_new_feature_rows(clone_R, clone_p, R_b2c, t_c_b, anchor, rho, f_an, pos, feat_off, obs_clone, obs_z, D)
    blocks, rs = [], []
    for j in range(Fk):
        sl = slice(feat_off[j], feat_off[j + 1])
        H, r = hy.feature_jacobian_ekf_new(clone_R, clone_p, R_b2c, t_c_b, obs_clone[sl], obs_z[sl], int(anchor[j]), D + j,
                                           D + Fk, [f_an[j, 0], f_an[j, 1], 1.0], rho[j], pos[j])
        blocks.append(H)
        rs.append(r)
    H_new, r_new = np.vstack(blocks), np.concatenate(rs)
    Hs, rs_ = hy.sparsify_new_features(H_new, r_new, Fk)
    rows = H_new.shape[0]
    Ho_ref, ro_ref = Hs[:rows - Fk, :D], rs_[:rows - Fk]
    H1_ref, H2_ref, r1_ref = Hs[rows - Fk:, :D], Hs[rows - Fk:, D:], rs_[rows - Fk:]
    assert out["H_o"].shape == Ho_ref.shape
    G, G_ref = out["H_o"].T @ out["H_o"], Ho_ref.T @ Ho_ref
    assert np.abs(G - G_ref).max() <= 1e-10 * np.abs(G_ref).max()
    b, b_ref = out["H_o"].T @ out["r_o"], Ho_ref.T @ ro_ref
    assert np.abs(b - b_ref).max() <= 1e-10 * max(np.abs(b_ref).max(), 1e-300)
    assert abs(out["r_o"] @ out["r_o"] - ro_ref @ ro_ref) <= 1e-10 * (ro_ref @ ro_ref)
    HH, HH_ref = out["H_1"] / out["h_2"][:, None], np.linalg.solve(H2_ref, H1_ref)
    assert np.abs(HH - HH_ref).max() <= 1e-10 * np.abs(HH_ref).max()
    assert np.abs(out["r_1"] / out["h_2"] - np.linalg.solve(H2_ref, r1_ref)).max() <= 1e-10 * max(np.abs(r1_ref / np.diag(H2_ref)).max(), 1e-300)
    np.testing.assert_allclose(np.abs(out["h_2"]), np.abs(np.diag(H2_ref)), rtol=1e-12)
    # delayed initialisation on top of a posterior
    A = rng.normal(0, 0.02, (D, D))
    P = A @ A.T + 1e-6 * np.eye(D)
    dx_leg = rng.normal(0, 1e-3, D)
    sigma2 = 6.4e-5
    dx_ref, P_ref = hy.delayed_initialization(P, dx_leg, H1_ref, H2_ref, r1_ref, sigma2)
    dx_new, P_aug = api.ekf_delayed_init(P, dx_leg, out["H_1"], out["h_2"], out["r_1"], sigma2)
    assert np.abs(dx_new - dx_ref).max() <= 1e-9 * max(np.abs(dx_ref).max(), 1e-300)
    assert np.abs(P_aug - P_ref).max() <= 1e-9 * np.abs(P_ref).max()
    assert np.abs(P_aug - P_aug.T).max() == 0.0


@pytest.mark.parametrize("seed,N,F,Fnew", [(9, 8, 4, 3), (10, 20, 24, 8)])
def test_hybrid_update_with_feature_states(seed, N, F, Fnew):
    """measurementUpdate_hybrid end to end at the stage level on a state with F inverse-depth features behind the
    clones: H_o = [MSCKF-like rows; gated rows of the F state features; feature-free rows of Fnew new features] through
    the dense whitened update with n = 6N + F window columns, then the delayed initialisation of the new features."""
    rng, R_b2c, t_c_b, clone_R, clone_p = _window(seed, N)
    D = 22 + 6 * N + F
    A = rng.normal(0, 0.02, (D, D))
    P = A @ A.T * 0.05 + 1e-6 * np.eye(D)
    P[15:22, :] = 0.0
    P[:, 15:22] = 0.0
    sigma2 = 6.4e-5
    # rows of the features already in the state (newest clone observes them)
    anchor, rho, f_an, pos, _, _, _ = _features(rng, R_b2c, t_c_b, clone_R, clone_p, F)
    anchor = np.minimum(anchor, N - 2).astype(np.int32)           # not anchored in the observing clone
    pos = np.array([hy.feature_position_from_anchor(clone_R[a], clone_p[a], R_b2c, t_c_b, [fx, fy, 1.0], r)
                    for a, r, (fx, fy) in zip(anchor, rho, f_an)])
    k = N - 1
    z = np.array([(lambda p_ck: p_ck[:2] / p_ck[2])(R_b2c @ clone_R[k].T @ (pos[f] - (clone_p[k] + clone_R[k] @ t_c_b)))
                  for f in range(F)]) + rng.normal(0, 0.004, (F, 2))
    rows_ekf = api.ekf_feature_rows(clone_R, clone_p, R_b2c, t_c_b, anchor, rho, f_an, pos, z, P, sigma2, 0.95)
    keep = np.repeat(rows_ekf["pass"].astype(bool), 2)
    H_ekf, r_ekf = rows_ekf["H"][keep], rows_ekf["r"][keep]
    assert H_ekf.shape[0] >= 2
    # new features (legacy dimension D): their feature-free rows join H_o
    a2, rho2, fan2, pos2, fo2, oc2, oz2 = _features(rng, R_b2c, t_c_b, clone_R, clone_p, Fnew + 6)
    ok = [f for f in range(Fnew + 6) if (oc2[fo2[f]:fo2[f + 1]] != a2[f]).sum() >= 2][:Fnew]
    fo, oc, oz = [0], [], []
    for f in ok:
        oc.extend(oc2[fo2[f]:fo2[f + 1]])
        oz.extend(oz2[fo2[f]:fo2[f + 1]])
        fo.append(len(oc))
    new = api.ekf_new_feature_rows(clone_R, clone_p, R_b2c, t_c_b, a2[ok], rho2[ok], fan2[ok], pos2[ok],
                                   np.array(fo, np.int32), np.array(oc, np.int32), np.array(oz), D)
    # a few generic clone-only rows standing in for the compressed MSCKF block
    H_m = np.zeros((3 * N, D))
    H_m[:, 22:22 + 6 * N] = rng.normal(0, 1.0, (3 * N, 6 * N))
    r_m = rng.normal(0, 0.01, 3 * N)
    H_o = np.vstack([H_m, H_ekf, new["H_o"]])
    r_o = np.concatenate([r_m, r_ekf, new["r_o"]])
    dx_ref, P_ref = hy.legacy_update(P, H_o, r_o, sigma2)
    dx, Pp = api.hybrid_update_dense(P, H_o, r_o, sigma2)
    assert np.abs(dx - dx_ref).max() <= 1e-9 * np.abs(dx_ref).max()
    assert np.abs(Pp - P_ref).max() <= 1e-9 * np.abs(P_ref).max()
    assert np.abs(Pp - Pp.T).max() == 0.0
    # ... and the new features enter the state
    H2 = np.diag(new["h_2"])
    dxn_ref, Pa_ref = hy.delayed_initialization(P_ref, dx_ref, new["H_1"], H2, new["r_1"], sigma2)
    dxn, Pa = api.ekf_delayed_init(Pp, dx, new["H_1"], new["h_2"], new["r_1"], sigma2)
    assert np.abs(dxn - dxn_ref).max() <= 1e-8 * max(np.abs(dxn_ref).max(), 1e-300)
    assert np.abs(Pa - Pa_ref).max() <= 1e-8 * np.abs(Pa_ref).max()
    assert Pa.shape == (D + len(ok), D + len(ok))


@pytest.mark.parametrize("N,E", [(5, 0), (12, 7), (29, 12)])
def test_augmentation_and_clone_removal_with_feature_block(N, E):
    rng = np.random.default_rng(40 + N)
    D = 22 + 6 * N + E
    A = rng.normal(0, 0.05, (D, D))
    P = A @ A.T + 1e-5 * np.eye(D)
    ref = hy.state_augmentation_cov(P, N)
    got = api.ekf_augment_cov(P, N)
    np.testing.assert_array_equal(got, ref)
    c = N // 2
    out = api.ekf_remove_clone_cov(got, N + 1, c)
    idx = np.r_[22 + 6 * c:28 + 6 * c]
    np.testing.assert_array_equal(out, np.delete(np.delete(got, idx, axis=0), idx, axis=1))
    assert out.shape == (D, D)


def test_hybrid_rows_against_the_committed_fixture():
    """The same entry points against tests/golden/hybrid_rows.npz -- no oracle import on this path."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hybrid_rows.npz"))
    N = g["clone_R"].shape[0]
    out = api.ekf_measurement_jacobians(g["clone_R"], g["clone_p"], g["R_b2c"], g["t_c_b"], g["anchor"], g["inv_depth"],
                                        g["f_an"], g["positions"], g["feat_off"], g["obs_clone"], g["obs_z"])
    for k in ("H_f", "H_a", "H_x", "H_e", "r"):
        assert np.abs(out[k] - g[k]).max() <= 1e-12 * max(np.abs(g[k]).max(), 1.0), k
    rows = api.ekf_feature_rows(g["clone_R"], g["clone_p"], g["R_b2c"], g["t_c_b"], g["anchor"], g["inv_depth"], g["f_an"],
                                g["positions"], g["z_cur"], g["P"], float(g["sigma2"]), 0.95)
    assert np.abs(rows["H"] - g["rows_H"]).max() <= 1e-12 * max(np.abs(g["rows_H"]).max(), 1.0)
    assert np.abs(rows["r"] - g["rows_r"]).max() <= 1e-12
    assert np.abs(rows["gamma"] - g["gamma"]).max() <= 1e-9 * np.abs(g["gamma"]).max()
    np.testing.assert_array_equal(rows["pass"].astype(bool), g["gamma"] < float(g["chi2"]))
    P_re, J_re = api.ekf_update_feature_cov(g["P"], g["clone_R"], g["clone_p"], g["R_b2c"], g["t_c_b"], int(g["re_fidx"]),
                                            int(g["re_old"]), int(g["re_new"]), g["positions"][int(g["re_fidx"])],
                                            float(g["re_rho_new"]))
    assert np.abs(J_re - g["J_re"]).max() <= 1e-12 * max(np.abs(g["J_re"]).max(), 1.0)
    assert np.abs(P_re - g["P_re"]).max() <= 1e-12 * np.abs(g["P_re"]).max()
    assert N == 8
