"""CPU: the oracle restatement of the hybrid EKF-SLAM feature Jacobians (oracle/hybrid.py)
against central differences of the measurement model -- the reference has no test for these functions, so this
is what pins the restatement (the reference checks its object Jacobians the same way,
src/tests/test_object_lm.cpp:493-545)."""
import numpy as np
import pytest

from oracle import hybrid as hy
from oracle import mathutils as mu


def _scene(seed):
    rng = np.random.default_rng(seed)
    R_a = mu.so3_exp(rng.normal(0, 0.4, 3))
    R_k = mu.so3_exp(rng.normal(0, 0.05, 3)) @ R_a
    t_a = rng.normal(0, 1.0, 3)
    t_k = t_a + rng.normal(0, 0.3, 3)
    R_b2c = mu.so3_exp(rng.normal(0, 0.8, 3))
    t_c_b = rng.normal(0, 0.1, 3)
    f_an = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.2, 0.2), 1.0])
    rho = 1.0 / rng.uniform(3.0, 15.0)
    z = rng.normal(0, 0.3, 2)
    return R_k, t_k, R_a, t_a, R_b2c, t_c_b, f_an, rho, z


def _predict(R_k, t_k, R_a, t_a, R_b2c, t_c_b, f_an, rho):
    p_w = hy.feature_position_from_anchor(R_a, t_a, R_b2c, t_c_b, f_an, rho)
    p_ck = R_b2c @ R_k.T @ (p_w - (t_k + R_k @ t_c_b))
    return p_ck[:2] / p_ck[2]


@pytest.mark.parametrize("seed", range(6))
def test_ekf_1didp_jacobians_match_central_differences(seed):
    R_k, t_k, R_a, t_a, R_b2c, t_c_b, f_an, rho, z = _scene(seed)
    p_w = hy.feature_position_from_anchor(R_a, t_a, R_b2c, t_c_b, f_an, rho)
    H_f, H_a, H_x, H_e, r = hy.measurement_jacobian_ekf_1didp(R_k, t_k, R_a, t_a, R_b2c, t_c_b, f_an, rho, p_w, z)
    np.testing.assert_allclose(r, z - _predict(R_k, t_k, R_a, t_a, R_b2c, t_c_b, f_an, rho), atol=1e-14)
    h = 1e-6

    def num(fun):
        return (fun(h) - fun(-h)) / (2 * h)

    # inverse depth
    np.testing.assert_allclose(H_f[:, 0], num(lambda e: _predict(R_k, t_k, R_a, t_a, R_b2c, t_c_b, f_an, rho + e)),
                               rtol=1e-6, atol=1e-8)
    for j in range(3):
        d = np.zeros(3)
        d[j] = 1.0
        # poses: R <- exp(dtheta) R (world-frame perturbation), p <- p + dp  (the LARVIO error state)
        np.testing.assert_allclose(H_x[:, j], num(lambda e: _predict(mu.so3_exp(e * d) @ R_k, t_k, R_a, t_a, R_b2c,
                                                                     t_c_b, f_an, rho)), rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(H_x[:, 3 + j], num(lambda e: _predict(R_k, t_k + e * d, R_a, t_a, R_b2c, t_c_b,
                                                                         f_an, rho)), rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(H_a[:, j], num(lambda e: _predict(R_k, t_k, mu.so3_exp(e * d) @ R_a, t_a, R_b2c,
                                                                     t_c_b, f_an, rho)), rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(H_a[:, 3 + j], num(lambda e: _predict(R_k, t_k, R_a, t_a + e * d, R_b2c, t_c_b,
                                                                         f_an, rho)), rtol=1e-6, atol=1e-8)
        # extrinsics: R_b2c <- R_b2c exp(-dphi)  (incrementState_IMUCam: R_ic <- R_ic dq^T, :4513-4518),
        # t_c_b <- t_c_b + dt
        np.testing.assert_allclose(H_e[:, j], num(lambda e: _predict(R_k, t_k, R_a, t_a, R_b2c @ mu.so3_exp(-e * d),
                                                                     t_c_b, f_an, rho)), rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(H_e[:, 3 + j], num(lambda e: _predict(R_k, t_k, R_a, t_a, R_b2c, t_c_b + e * d,
                                                                         f_an, rho)), rtol=1e-6, atol=1e-8)


def test_anchor_observation_is_zeroed():
    R_k, t_k, R_a, t_a, R_b2c, t_c_b, f_an, rho, z = _scene(11)
    p_w = hy.feature_position_from_anchor(R_a, t_a, R_b2c, t_c_b, f_an, rho)
    out = hy.measurement_jacobian_ekf_1didp(R_a, t_a, R_a, t_a, R_b2c, t_c_b, f_an, rho, p_w, z, same_state=True)
    assert all(np.all(o == 0) for o in out)


def test_stacked_rows_and_gate():
    rng = np.random.default_rng(5)
    N, E = 6, 3
    R_b2c, t_c_b = mu.so3_exp(rng.normal(0, 0.8, 3)), rng.normal(0, 0.1, 3)
    clone_R = [mu.so3_exp(rng.normal(0, 0.05, 3)) for _ in range(N)]
    clone_p = [np.array([0.2 * i, 0.0, 0.0]) + rng.normal(0, 0.02, 3) for i in range(N)]
    a_idx, k_idx, feat_idx = 1, N - 1, 2
    f_an, rho = np.array([0.1, -0.05, 1.0]), 0.2
    p_w = hy.feature_position_from_anchor(clone_R[a_idx], clone_p[a_idx], R_b2c, t_c_b, f_an, rho)
    z = _predict(clone_R[k_idx], clone_p[k_idx], clone_R[a_idx], clone_p[a_idx], R_b2c, t_c_b, f_an, rho) + 1e-3
    H, r = hy.feature_jacobian_ekf(clone_R, clone_p, R_b2c, t_c_b, k_idx, a_idx, feat_idx, E, f_an, rho, p_w, z)
    D = 22 + 6 * N + E
    assert H.shape == (2, D)
    nz = np.flatnonzero(np.abs(H).sum(axis=0))
    expect = set(range(15, 21)) | set(range(22 + 6 * a_idx, 28 + 6 * a_idx)) | set(range(22 + 6 * k_idx, 28 + 6 * k_idx))
    expect.add(22 + 6 * N + feat_idx)
    assert set(nz) <= expect and (22 + 6 * N + feat_idx) in nz
    np.testing.assert_allclose(r, [1e-3, 1e-3], atol=1e-12)
    A = rng.normal(0, 0.05, (D, D))
    P = A @ A.T + 1e-4 * np.eye(D)
    g, ok = hy.gate_ekf_row(H, r, P, 6.4e-5, mu.chi2_table(0.95)[2])
    assert g > 0 and ok
    # a new feature: rows for every observing clone but the anchor
    obs_clone = [1, 2, 3, 5]
    obs_z = [_predict(clone_R[c], clone_p[c], clone_R[a_idx], clone_p[a_idx], R_b2c, t_c_b, f_an, rho) for c in obs_clone]
    Hn, rn = hy.feature_jacobian_ekf_new(clone_R, clone_p, R_b2c, t_c_b, obs_clone, obs_z, a_idx, D, D + 1, f_an, rho, p_w)
    assert Hn.shape == (6, D + 1) and np.abs(rn).max() < 1e-12
    assert np.all(np.abs(Hn[:, D]) > 0)


def _inv_depth_in(R_anchor, t_anchor, R_b2c, t_c_b, p_w):
    p_c = R_b2c @ R_anchor.T @ (p_w - (t_anchor + R_anchor @ t_c_b))
    return 1.0 / p_c[2]


@pytest.mark.parametrize("seed", range(4))
def test_reanchor_jacobian_matches_central_differences(seed):
    """updateFeatureCov_1didp: rho_new as a function of (rho_old, old anchor pose, new anchor pose, extrinsics)."""
    R_new, t_new, R_old, t_old, R_b2c, t_c_b, f_an, rho_old, _ = _scene(20 + seed)

    def rho_new(R_o=R_old, t_o=t_old, R_n=R_new, t_n=t_new, Rbc=R_b2c, tcb=t_c_b, rho=rho_old):
        p_w = hy.feature_position_from_anchor(R_o, t_o, Rbc, tcb, f_an, rho)
        return _inv_depth_in(R_n, t_n, Rbc, tcb, p_w)

    p_w = hy.feature_position_from_anchor(R_old, t_old, R_b2c, t_c_b, f_an, rho_old)
    H_f, H_old, H_new, H_e = hy.reanchor_jacobian(R_old, t_old, R_new, t_new, R_b2c, t_c_b, p_w, rho_new())
    h = 1e-6

    def num(fun):
        return (fun(h) - fun(-h)) / (2 * h)

    np.testing.assert_allclose(H_f, num(lambda e: rho_new(rho=rho_old + e)), rtol=1e-6, atol=1e-9)
    for j in range(3):
        d = np.zeros(3)
        d[j] = 1.0
        np.testing.assert_allclose(H_old[j], num(lambda e: rho_new(R_o=mu.so3_exp(e * d) @ R_old)), rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(H_old[3 + j], num(lambda e: rho_new(t_o=t_old + e * d)), rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(H_new[j], num(lambda e: rho_new(R_n=mu.so3_exp(e * d) @ R_new)), rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(H_new[3 + j], num(lambda e: rho_new(t_n=t_new + e * d)), rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(H_e[j], num(lambda e: rho_new(Rbc=R_b2c @ mu.so3_exp(-e * d))), rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(H_e[3 + j], num(lambda e: rho_new(tcb=t_c_b + e * d)), rtol=1e-6, atol=1e-9)


def test_update_feature_cov_is_a_congruence():
    rng = np.random.default_rng(9)
    N, E = 5, 3
    R_b2c, t_c_b = mu.so3_exp(rng.normal(0, 0.8, 3)), rng.normal(0, 0.1, 3)
    clone_R = [mu.so3_exp(rng.normal(0, 0.05, 3)) for _ in range(N)]
    clone_p = [np.array([0.2 * i, 0.0, 0.0]) + rng.normal(0, 0.02, 3) for i in range(N)]
    f_an, rho = np.array([0.1, -0.05, 1.0]), 0.15
    old, new, fidx = 0, 3, 1
    p_w = hy.feature_position_from_anchor(clone_R[old], clone_p[old], R_b2c, t_c_b, f_an, rho)
    rho_n = _inv_depth_in(clone_R[new], clone_p[new], R_b2c, t_c_b, p_w)
    D = 22 + 6 * N + E
    A = rng.normal(0, 0.05, (D, D))
    P = A @ A.T + 1e-4 * np.eye(D)
    Pn, J = hy.update_feature_cov_1didp(P, N, fidx, old, new, clone_R, clone_p, R_b2c, t_c_b, p_w, rho_n)
    T = np.eye(D)
    T[22 + 6 * N + fidx] = J[0]
    np.testing.assert_allclose(Pn, T @ P @ T.T, rtol=1e-12, atol=1e-15)      # P' = T P T^T: still symmetric PSD
    assert np.linalg.eigvalsh(Pn).min() > 0


def test_new_feature_rows_decouple_and_initialise():
    """H3 + the delayed initialisation: for 1-D inverse-depth features the feature part of the stacked new rows has
    one column per feature with disjoint row supports, so H_2 is diagonal, and the initialised inverse depths
    reproduce the measurements (noise-free case: dx_new closes the residual exactly to first order)."""
    rng = np.random.default_rng(3)
    N = 7
    R_b2c, t_c_b = mu.so3_exp(rng.normal(0, 0.8, 3)), rng.normal(0, 0.1, 3)
    clone_R = [mu.so3_exp(rng.normal(0, 0.05, 3)) for _ in range(N)]
    clone_p = [np.array([0.25 * i, 0.0, 0.0]) + rng.normal(0, 0.02, 3) for i in range(N)]
    D = 22 + 6 * N
    feats = [(0, [0, 1, 2, 4], 0.12), (3, [1, 3, 5, 6], 0.2), (6, [2, 4, 5, 6], 0.08)]   # (anchor, observers, rho)
    sz = len(feats)
    blocks, rs = [], []
    for j, (a, obs, rho) in enumerate(feats):
        f_an = np.array([0.05 * j, -0.03 * j, 1.0])
        rho_guess = rho * 1.05                                   # linearisation point off by 5 %
        p_true = hy.feature_position_from_anchor(clone_R[a], clone_p[a], R_b2c, t_c_b, f_an, rho)
        p_lin = hy.feature_position_from_anchor(clone_R[a], clone_p[a], R_b2c, t_c_b, f_an, rho_guess)
        zs = []
        for c in obs:
            p_ck = R_b2c @ clone_R[c].T @ (p_true - (clone_p[c] + clone_R[c] @ t_c_b))
            zs.append(p_ck[:2] / p_ck[2])
        H, r = hy.feature_jacobian_ekf_new(clone_R, clone_p, R_b2c, t_c_b, obs, zs, a, D + j, D + sz, f_an, rho_guess, p_lin)
        blocks.append(H)
        rs.append(r)
    H_new, r_new = np.vstack(blocks), np.concatenate(rs)
    Hs, rs_ = hy.sparsify_new_features(H_new, r_new, sz)
    rows = H_new.shape[0]
    assert np.abs(Hs[:rows - sz, D:]).max() < 1e-12               # the nullspace rows lost their feature part
    H_1, H_2, r_1 = Hs[rows - sz:, :D], Hs[rows - sz:, D:], rs_[rows - sz:]
    assert np.abs(H_2 - np.diag(np.diag(H_2))).max() < 1e-12      # diagonal: the reference's ldlt() solve is valid
    A = rng.normal(0, 0.01, (D, D))
    P = A @ A.T + 1e-6 * np.eye(D)
    dx_new, P_aug = hy.delayed_initialization(P, np.zeros(D), H_1, H_2, r_1, 6.4e-5)
    for j, (a, obs, rho) in enumerate(feats):
        assert abs((rho * 1.05 + dx_new[j]) - rho) < 2e-3 * rho   # one Gauss-Newton step from 5 % off
    assert np.linalg.eigvalsh(P_aug).min() > 0 and np.abs(P_aug - P_aug.T).max() == 0
    # invariance under the basis: any orthogonal mixing of the bottom rows gives the same initialisation
    Qm = np.linalg.qr(rng.normal(size=(sz, sz)))[0]
    dx2, P2 = hy.delayed_initialization(P, np.zeros(D), Qm @ H_1, Qm @ H_2, Qm @ r_1, 6.4e-5)
    np.testing.assert_allclose(dx2, dx_new, rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(P2, P_aug, rtol=1e-9, atol=1e-18)


def test_augmentation_with_feature_block_reduces_to_the_msckf_case():
    """oracle/hybrid.state_augmentation_cov with E = 0 is what oracle/filter.OracleVIO.stateAugmentation does; with
    E > 0 it is the same congruence with the new block moved in front of the feature block."""
    rng = np.random.default_rng(2)
    N, E = 4, 3
    D = 22 + 6 * N + E
    A = rng.normal(0, 0.1, (D, D))
    P = A @ A.T
    out = hy.state_augmentation_cov(P, N)
    T = np.zeros((D + 6, D))
    pose = 22 + 6 * N
    T[:pose, :pose] = np.eye(pose)
    T[pose:pose + 3, 0:3] = np.eye(3)
    T[pose + 3:pose + 6, 6:9] = np.eye(3)
    T[pose + 6:, pose:] = np.eye(E)
    np.testing.assert_allclose(out, T @ P @ T.T, rtol=1e-13, atol=1e-16)
    P0 = P[:pose, :pose]
    out0 = hy.state_augmentation_cov(P0, N)
    J = np.zeros((6, pose))
    J[0:3, 0:3] = np.eye(3)
    J[3:6, 6:9] = np.eye(3)
    ref0 = np.block([[P0, (J @ P0).T], [J @ P0, J @ P0 @ J.T]])
    np.testing.assert_allclose(out0, (ref0 + ref0.T) / 2, rtol=1e-13, atol=1e-16)


def test_oracle_reproduces_the_committed_hybrid_fixture():
    """tests/golden/hybrid_rows.npz (made by tests/golden/make_hybrid_golden.py) pins the hybrid restatement."""
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_hybrid_golden", os.path.join(here, "golden", "make_hybrid_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    now = mod.build()
    ref = np.load(os.path.join(here, "golden", "hybrid_rows.npz"))
    for k in ref.files:
        np.testing.assert_allclose(np.asarray(now[k], dtype=float), np.asarray(ref[k], dtype=float), rtol=1e-12, atol=1e-14,
                                   err_msg=k)
