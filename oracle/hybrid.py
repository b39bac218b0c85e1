"""Oracle (test infrastructure, CPU restatement) for the hybrid EKF-SLAM feature rows of SURVEY 8a:

  H1  measurementJacobian_ekf_1didp            reference src/orcvio.cpp:1356-1478
  H2  featureJacobian_ekf / featureJacobian_ekf_new   reference src/orcvio.cpp:1575-1651 / 1481-1572

restricted to what every shipped yaml uses: feature_idp_dim == 1 (1-D inverse depth), use_schmidt == 0 (the
anchor is a clone of the window, never a nuisance state), if_FEJ == 0, estimate_td == 0.

Parity unpinned by the reference (it has no test for these functions): pinned here by central differences of the
measurement model the formulas differentiate (tests/test_oracle_hybrid_cpu.py) and by the CUDA kernel agreeing
with this restatement (tests/test_gpu_hybrid.py).  Only tests/ may import this module.
"""
import numpy as np

from . import mathutils as mu

LEG_DIM = 22


def measurement_jacobian_ekf_1didp(R_bk2w, t_bk_w, R_ba2w, t_ba_w, R_b2c, t_c_b, f_an, inv_depth, p_w, z,
                                   same_state=False):
    """:1356-1478.  Clone k observes a feature anchored in clone a with inverse depth `inv_depth` along the
    anchor-frame bearing f_an = (x, y, 1); p_w is the feature's world position as stored in the map server
    (the reference reads feature.position, it does not recompute it).  Returns H_f (2x1), H_a (2x6, anchor
    pose), H_x (2x6, pose of clone k), H_e (2x6, extrinsics), r (2).  `same_state`: state_id == id_anchor,
    for which the reference returns zeros (:1433-1441)."""
    R_bk2w, R_ba2w, R_b2c = (np.asarray(a, dtype=float).reshape(3, 3) for a in (R_bk2w, R_ba2w, R_b2c))
    t_bk_w, t_ba_w, t_c_b, f_an, p_w = (np.asarray(a, dtype=float).reshape(3) for a in (t_bk_w, t_ba_w, t_c_b, f_an, p_w))
    z = np.asarray(z, dtype=float).reshape(2)
    R_w2bk = R_bk2w.T
    R_w2ck = R_b2c @ R_w2bk
    t_ck_w = t_bk_w + R_bk2w @ t_c_b
    R_w2ba = R_ba2w.T
    R_w2ca = R_b2c @ R_w2ba
    p_ca = np.array([f_an[0] / inv_depth, f_an[1] / inv_depth, 1.0 / inv_depth])
    p_ck = R_w2ck @ (p_w - t_ck_w)
    r = z - np.array([p_ck[0] / p_ck[2], p_ck[1] / p_ck[2]])
    if same_state:
        return np.zeros((2, 1)), np.zeros((2, 6)), np.zeros((2, 6)), np.zeros((2, 6)), np.zeros(2)
    J_k = np.zeros((2, 3))
    J_k[0, 0] = 1 / p_ck[2]
    J_k[1, 1] = 1 / p_ck[2]
    J_k[0, 2] = -p_ck[0] / (p_ck[2] * p_ck[2])
    J_k[1, 2] = -p_ck[1] / (p_ck[2] * p_ck[2])
    J_d = R_w2ck @ R_w2ca.T @ f_an
    p_baf_w = p_w - t_ba_w
    p_bkf_w = p_w - t_bk_w
    J_xa = np.zeros((3, 6))
    J_xa[:, :3] = -R_w2ck @ mu.skew(p_baf_w)
    J_xa[:, 3:] = R_w2ck
    J_xk = np.zeros((3, 6))
    J_xk[:, :3] = R_w2ck @ mu.skew(p_bkf_w)
    J_xk[:, 3:] = -R_w2ck
    J_e = np.zeros((3, 6))
    SkewMx = mu.skew(R_w2bk @ p_bkf_w - t_c_b)
    Mx = R_w2bk @ R_w2ba.T @ mu.skew(R_b2c.T @ p_ca)
    J_e[:, :3] = R_b2c @ (SkewMx - Mx)
    J_e[:, 3:] = R_b2c @ (R_w2bk @ R_w2ba.T - np.eye(3))
    J_rho = -1.0 / (inv_depth * inv_depth)
    H_f = (J_k @ J_d * J_rho).reshape(2, 1)
    return H_f, J_k @ J_xa, J_k @ J_xk, J_k @ J_e, r


def feature_position_from_anchor(R_ba2w, t_ba_w, R_b2c, t_c_b, f_an, inv_depth):
    """World position of an inverse-depth feature: p_w = R_ca2w p_ca + t_ca_w (measurementUpdate_hybrid
    :1866-1877 with orientation_cam = R_b2w R_b2c^T, position_cam = t_b_w + R_b2w t_c_b)."""
    R_ba2w, R_b2c = np.asarray(R_ba2w, dtype=float).reshape(3, 3), np.asarray(R_b2c, dtype=float).reshape(3, 3)
    p_ca = np.array([f_an[0] / inv_depth, f_an[1] / inv_depth, 1.0 / inv_depth])
    return R_ba2w @ R_b2c.T @ p_ca + (np.asarray(t_ba_w, dtype=float) + R_ba2w @ np.asarray(t_c_b, dtype=float))


def feature_jacobian_ekf(clone_R, clone_p, R_b2c, t_c_b, k_idx, a_idx, feat_idx, n_feat_states, f_an, inv_depth,
                         p_w, z):
    """:1575-1651.  One EKF-SLAM feature already in the state, observed by the newest state (clone k_idx):
    the 2 x D row block with H_f at the feature's column LEG + 6N + feat_idx, H_a at the anchor clone's block,
    H_x at the observing clone's block, H_e at columns 15..20.  D = LEG + 6N + n_feat_states."""
    N = len(clone_R)
    D = LEG_DIM + 6 * N + n_feat_states
    H_f, H_a, H_x, H_e, r = measurement_jacobian_ekf_1didp(clone_R[k_idx], clone_p[k_idx], clone_R[a_idx],
                                                           clone_p[a_idx], R_b2c, t_c_b, f_an, inv_depth, p_w, z,
                                                           same_state=(k_idx == a_idx))
    H = np.zeros((2, D))
    H[:, LEG_DIM + 6 * N + feat_idx] = H_f[:, 0]
    H[:, LEG_DIM + 6 * a_idx:LEG_DIM + 6 * a_idx + 6] = H_a
    H[:, LEG_DIM + 6 * k_idx:LEG_DIM + 6 * k_idx + 6] = H_x       # written after H_a like the reference (:1644-1645)
    H[:, 15:21] = H_e
    return H, r


def feature_jacobian_ekf_new(clone_R, clone_p, R_b2c, t_c_b, obs_clone, obs_z, a_idx, feat_col, n_cols, f_an,
                             inv_depth, p_w):
    """:1481-1572.  A feature about to enter the state: rows for every observing clone except the anchor
    (1-D inverse depth: the anchor observation carries no information, :1497-1498), with H_f at column `feat_col`
    of an n_cols-wide block."""
    rows = [(c, z) for c, z in zip(obs_clone, obs_z) if c != a_idx]
    H = np.zeros((2 * len(rows), n_cols))
    r = np.zeros(2 * len(rows))
    for i, (c, z) in enumerate(rows):
        H_f, H_a, H_x, H_e, r_i = measurement_jacobian_ekf_1didp(clone_R[c], clone_p[c], clone_R[a_idx], clone_p[a_idx],
                                                                 R_b2c, t_c_b, f_an, inv_depth, p_w, z)
        H[2 * i:2 * i + 2, feat_col] = H_f[:, 0]
        H[2 * i:2 * i + 2, LEG_DIM + 6 * a_idx:LEG_DIM + 6 * a_idx + 6] = H_a
        H[2 * i:2 * i + 2, LEG_DIM + 6 * c:LEG_DIM + 6 * c + 6] = H_x
        H[2 * i:2 * i + 2, 15:21] = H_e
        r[2 * i:2 * i + 2] = r_i
    return H, r


def gate_ekf_row(H, r, P, sigma2, chi2_dof2):
    """gatingTestFeature(H_xj, r_j, 2) (:1953-1976) on one 2 x D block."""
    S = H @ P @ H.T + sigma2 * np.eye(2)
    gamma = float(r @ np.linalg.solve(S, r))
    return gamma, gamma < chi2_dof2


def reanchor_jacobian(R_old, t_old, R_new, t_new, R_b2c, t_c_b, p_w, inv_depth_new):
    """updateFeatureCov_1didp (:3611-3699): d(rho_new) with respect to (rho_old, old anchor pose, new anchor pose,
    extrinsics) when a 1-D inverse-depth feature moves its anchor from clone `old` to clone `new`.  p_w is the
    feature's world position, inv_depth_new the inverse depth already re-expressed in the new anchor (the caller
    sets feature.invDepth before the call, :2820-2860).  Returns (H_f, H_old (6), H_new (6), H_e (6))."""
    R_old, R_new, R_b2c = (np.asarray(a, dtype=float).reshape(3, 3) for a in (R_old, R_new, R_b2c))
    t_old, t_new, t_c_b, p_w = (np.asarray(a, dtype=float).reshape(3) for a in (t_old, t_new, t_c_b, p_w))
    R_c2w_old = R_old @ R_b2c.T
    t_c_w_old = t_old + R_old @ t_c_b
    p_old = np.linalg.solve(R_c2w_old, p_w - t_c_w_old)       # R_c2w_old.inverse() * (...)
    inv_old = 1 / p_old[2]
    f_old = np.array([p_old[0] / p_old[2], p_old[1] / p_old[2], 1.0])
    R_w2b_new = R_new.T
    R_c2w_new = R_new @ R_b2c.T
    R_w2c_new = R_c2w_new.T
    p_bf_old = p_w - t_old
    p_bf_new = p_w - t_new
    J_rho_d_new = -inv_depth_new * inv_depth_new
    J_d = (R_w2c_new @ R_c2w_old @ f_old)[2]
    J_theta_old = (-R_w2c_new @ mu.skew(p_bf_old))[2]
    J_p_old = R_w2c_new[2]
    J_theta_new = (R_w2c_new @ mu.skew(p_bf_new))[2]
    J_p_new = -R_w2c_new[2]
    SkewMx = mu.skew(R_w2b_new @ p_bf_new - t_c_b)
    Mx = R_w2b_new @ R_old @ mu.skew(R_b2c.T @ p_old)
    J_e_theta = (R_b2c @ (SkewMx - Mx))[2]
    J_e_p = (R_b2c @ (R_w2b_new @ R_old - np.eye(3)))[2]
    J_d_rho_old = -1 / (inv_old * inv_old)
    H_f = J_rho_d_new * J_d * J_d_rho_old
    return (H_f, J_rho_d_new * np.concatenate([J_theta_old, J_p_old]),
            J_rho_d_new * np.concatenate([J_theta_new, J_p_new]), J_rho_d_new * np.concatenate([J_e_theta, J_e_p]))


def update_feature_cov_1didp(P, n_clones, feat_idx, old_idx, new_idx, clone_R, clone_p, R_b2c, t_c_b, p_w,
                             inv_depth_new):
    """updateFeatureCov_1didp (:3611-3773), use_schmidt == 0: the feature's row / column of P is replaced by J P
    (J P J^T on the diagonal), J the 1 x D Jacobian above; P is then symmetrised."""
    P = np.array(P, dtype=float)
    D = P.shape[0]
    H_f, H_old, H_new, H_e = reanchor_jacobian(clone_R[old_idx], clone_p[old_idx], clone_R[new_idx], clone_p[new_idx],
                                               R_b2c, t_c_b, p_w, inv_depth_new)
    c = LEG_DIM + 6 * n_clones + feat_idx
    J = np.zeros((1, D))
    J[0, c] = H_f
    J[0, LEG_DIM + 6 * old_idx:LEG_DIM + 6 * old_idx + 6] = H_old
    J[0, LEG_DIM + 6 * new_idx:LEG_DIM + 6 * new_idx + 6] = H_new      # written after the old block (:3714-3717)
    J[0, 15:21] = H_e
    Pfleg = J @ P
    Pff = Pfleg @ J.T
    out = P.copy()
    out[c, :] = Pfleg[0]
    out[:, c] = Pfleg[0]
    out[c, c] = Pff[0, 0]
    return (out + out.T) / 2.0, J


def sparsify_new_features(H_ekf_new, r_ekf_new, sz_new):
    """New-feature sparsification of removeLostFeatures (:2413-2443): W = [V U] with V the left nullspace of
    H_f_new (full SVD) and U its column space (QR); returns W^T H, W^T r -- the last sz_new rows are (H_1 | H_2),
    r_1 of measurementUpdate_hybrid, the rows above have a zero feature part."""
    H = np.asarray(H_ekf_new, dtype=float)
    r = np.asarray(r_ekf_new, dtype=float)
    H_f = H[:, H.shape[1] - sz_new:]
    Us = np.linalg.svd(H_f, full_matrices=True)[0]
    V = Us[:, sz_new:]
    Q = np.linalg.qr(H_f, mode="complete")[0]
    W = np.concatenate([V, Q[:, :sz_new]], axis=1)
    return W.T @ H, W.T @ r


def delayed_initialization(P_post, dx_leg, H_1, H_2, r_1, sigma2):
    """The new-state part of measurementUpdate_hybrid (:1823-1832, 1903-1941; use_schmidt == 0): given the
    posterior of the legacy state and the rows (H_1 | H_2), r_1 that carry the new features,
        HH = H_2^-1 H_1,  dx_new = -HH dx_leg + H_2^-1 r_1,
        P_aug = [[P, -P HH^T], [-HH P, HH P HH^T + sigma^2 (H_2^T H_2)^-1]]  (then symmetrised).
    H_2 is diagonal for 1-D inverse-depth features (each new feature's column of H_f has its own rows), which is
    what makes the reference's H_2.ldlt() -- a factorisation of the lower triangle -- a valid solve."""
    P = np.asarray(P_post, dtype=float)
    H_1, H_2 = np.asarray(H_1, dtype=float), np.asarray(H_2, dtype=float)
    HH = np.linalg.solve(H_2, H_1)
    dx_new = -HH @ dx_leg + np.linalg.solve(H_2, r_1)
    nHHP = -HH @ P
    P22 = -nHHP @ HH.T + sigma2 * np.linalg.inv(H_2.T @ H_2)
    D, sz = P.shape[0], H_2.shape[0]
    out = np.zeros((D + sz, D + sz))
    out[:D, :D] = P
    out[D:, :D] = nHHP
    out[:D, D:] = nHHP.T
    out[D:, D:] = P22
    return dx_new, (out + out.T) / 2.0


def legacy_update(P, H_o, r_o, sigma2):
    """Legacy-state part of measurementUpdate_hybrid (:1808-1820, 1884-1901; no Schmidt):
    S = H P H^T + sigma^2 I, K^T = S^-1 H P, dx = K r, P <- (I - K H) P, symmetrised."""
    P = np.asarray(P, dtype=float)
    H = np.asarray(H_o, dtype=float)
    S = H @ P @ H.T + sigma2 * np.eye(H.shape[0])
    K = np.linalg.solve(S, H @ P).T
    dx = K @ np.asarray(r_o, dtype=float)
    Pn = (np.eye(P.shape[0]) - K @ H) @ P
    return dx, (Pn + Pn.T) / 2.0


def state_augmentation_cov(P, n_clones):
    """stateAugmentation, covariance part (:963-1010), with E = D - 22 - 6 n_clones feature states behind the clones:
    P12 = J P, P11 = P12 J^T, the new 6 x 6 block is inserted before the feature block, then (P + P^T)/2."""
    P = np.asarray(P, dtype=float)
    D = P.shape[0]
    pose = LEG_DIM + 6 * n_clones
    rest = D - pose
    J = np.zeros((6, D))
    J[0:3, 0:3] = np.eye(3)
    J[3:6, 6:9] = np.eye(3)
    P12 = J @ P
    P11 = P12 @ J.T
    out = np.zeros((D + 6, D + 6))
    out[:pose, :pose] = P[:pose, :pose]
    out[pose + 6:, :pose] = P[pose:, :pose]
    out[:pose, pose + 6:] = P[:pose, pose:]
    out[pose + 6:, pose + 6:] = P[pose:, pose:]
    out[pose:pose + 6, pose:pose + 6] = P11
    out[pose:pose + 6, :pose] = P12[:, :pose]
    out[:pose, pose:pose + 6] = P12[:, :pose].T
    out[pose:pose + 6, pose + 6:] = P12[:, pose:]
    out[pose + 6:, pose:pose + 6] = P12[:, pose:].T
    assert rest >= 0
    return (out + out.T) / 2.0
