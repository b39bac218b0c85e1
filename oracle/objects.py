"""CPU oracle -- object keypoint / bounding-box residual model, stage 3
(TEST INFRASTRUCTURE, not product code).

NumPy restatement of the residual + Jacobian functors evaluated at the object-LM
optimum to produce (r, H_f, H_c) for the filter:
  O1  CameraLM::ErrorFeatureQuadric   src/obj/ObjectResJacCam.cpp:153-282
  O2  CameraLM::ErrorBBoxQuadric      src/obj/ObjectResJacCam.cpp:308-519
  O3  CameraLM::operator()/df, get_valid_camera_pose_mat   :521-604
  O4  ObjectLM::ErrorFeatureQuadric / ErrorBBoxQuadric (object-state Jacobians)
      src/obj/ObjectLM.cpp:272-371, :443-632
helpers bbox2poly / poly2lineh / ellipse_from_shape  src/obj/ObjectLM.cpp:380-414,
projection Jacobians include/orcvio/utils/se3_ops.hpp:325-453.

parity: PINNED by the reference's golden vectors test_error_feature_quadric.h5 and
test_error_bbox_quadric.h5 (src/tests/test_object_lm.cpp:90-202, 482-584), committed
as tests/golden/*.npz by tests/golden/make_golden.py.
"""
import math
import numpy as np

from . import mathutils as mu

PS = np.hstack([np.eye(3), np.zeros((3, 1))])    # ps_puline_s, se3_ops.hpp:422-423


def valid_indices(zs):
    """filter_valid_indices, ObjectLM.cpp:190-198 (rows with all-finite entries)."""
    return [i for i in range(zs.shape[0]) if np.all(np.isfinite(zs[i]))]


def bbox2poly(bbox):
    """ObjectLM.cpp:380-392."""
    xmin, ymin, xmax, ymax = bbox
    return np.array([[xmin, ymin], [xmax, ymin], [xmax, ymax], [xmin, ymax]], dtype=float)


def poly2lineh(points):
    """ObjectLM.cpp:394-405."""
    n = points.shape[0]
    out = np.zeros((n, 3))
    for i in range(n):
        a = np.array([points[i, 0], points[i, 1], 1.0])
        b = np.array([points[(i + 1) % n, 0], points[(i + 1) % n, 1], 1.0])
        out[i] = np.cross(a, b)
    return out


def ellipse_from_shape(v):
    """ObjectLM.cpp:407-414: diag(v^2, -1)."""
    return np.diag([v[0] * v[0], v[1] * v[1], v[2] * v[2], -1.0])


def project_object_points(P, wTo, pts_h):
    """se3_ops.hpp:349-355: pts_h (n x 4) -> (n x 2)."""
    uvh = P @ (wTo @ pts_h.T)
    return (uvh[:2] / uvh[2]).T


# ---------------------------------------------------------------- O1: keypoints, camera pose
def kp_residual(cTw, wTo, kps_h, zs):
    """CameraLM::ErrorFeatureQuadric::operator() per frame, ObjectResJacCam.cpp:153-176
    (identical to ObjectLM.cpp:272-295): [(u1,v1),(u2,v2),...] over valid keypoints."""
    idx = valid_indices(zs)
    P = cTw[:3, :]
    uv = project_object_points(P, wTo, kps_h[idx])
    return (uv - zs[idx]).reshape(-1)


def kp_jac_camera(cTw, wTo, kps_h, zs, left):
    """project_object_points_df_camera, se3_ops.hpp:411-453 (2k x 6)."""
    idx = valid_indices(zs)
    P = cTw[:3, :]
    J = np.zeros((2 * len(idx), 6))
    for r, i in enumerate(idx):
        X = kps_h[i]
        dpi = mu.project_image_df(P @ wTo @ X)
        if left:
            jac = -1 * dpi @ PS @ cTw @ mu.odot(wTo @ X)
        else:
            jac = -1 * dpi @ PS @ mu.odot(cTw @ wTo @ X)
        J[2 * r:2 * r + 2] = jac
    return J


# ---------------------------------------------------------------- O4: keypoints, object state
def kp_jac_object(cTw, wTo, kps_h, zs, left):
    """ObjectLM::ErrorFeatureQuadric::df per frame, ObjectLM.cpp:318-349:
    (2k x (9+3K)) = [pose 6 | shape 3 (zero) | keypoints 3K]."""
    K = kps_h.shape[0]
    idx = valid_indices(zs)
    P = cTw[:3, :]
    J = np.zeros((2 * len(idx), 9 + 3 * K))
    for r, i in enumerate(idx):
        X = kps_h[i]
        dpi = mu.project_image_df(P @ wTo @ X)
        if left:
            J[2 * r:2 * r + 2, 0:6] = dpi @ P @ mu.odot(wTo @ X)      # se3_ops.hpp:386
        else:
            J[2 * r:2 * r + 2, 0:6] = dpi @ P @ wTo @ mu.odot(X)      # se3_ops.hpp:391
        J[2 * r:2 * r + 2, 9 + 3 * i:12 + 3 * i] = dpi @ P @ wTo[:, :3]  # ObjectLM.cpp:341-343
    return J


# ---------------------------------------------------------------- O2: bbox
def _new_bbox_terms(Pm, line, U_square):
    uline_b = Pm.T @ line
    b = uline_b[:3]
    b_norm = math.sqrt(float(b @ b))
    dist = uline_b[3]
    sqrt_bU2b = math.sqrt(float(b @ U_square @ b))
    sign = 1.0 if dist > 0 else -1.0
    return uline_b, b, b_norm, dist, sqrt_bU2b, sign


def bbox_residual(cTw, wTo, shape, bbox, new_residual):
    """ErrorBBoxQuadric::operator() per frame, ObjectResJacCam.cpp:308-349
    (== ObjectLM.cpp:443-484)."""
    Qi = ellipse_from_shape(shape)
    P = (cTw @ wTo)[:3, :]
    lines = poly2lineh(bbox2poly(bbox))
    if not new_residual:
        Ci = P @ Qi @ P.T
        return np.sum((lines @ Ci) * lines, axis=1)
    out = np.zeros(4)
    U2 = Qi[:3, :3]
    for i in range(4):
        _, _, b_norm, dist, sq, sign = _new_bbox_terms(P, lines[i], U2)
        out[i] = (dist - sign * sq) / b_norm
    return out


def _new_bbox_chain(uline_b, b_norm, sqrt_bU2b, sign, Qi):
    term1a = np.array([[0.0, 0.0, 0.0, 1.0]])
    term2a = Qi.copy()
    term2a[3, 3] = 0
    p_be_p_ulinea = term1a - sign * (uline_b[None, :] @ term2a) / sqrt_bU2b
    term1b = np.eye(4) / b_norm
    term2b = np.eye(4)
    term2b[3, 3] = 0
    p_ulinea_ulineb = term1b - np.outer(uline_b, uline_b) @ term2b / b_norm ** 3
    return p_be_p_ulinea @ p_ulinea_ulineb


def bbox_jac_camera(cTw, wTo, shape, bbox, left, new_residual):
    """CameraLM::ErrorBBoxQuadric::df per frame, ObjectResJacCam.cpp:396-499 (4 x 6)."""
    Qi = ellipse_from_shape(shape)
    P = cTw[:3, :]
    P_prime = np.eye(4)[:3, :]
    lines = poly2lineh(bbox2poly(bbox))
    J = np.zeros((4, 6))
    for i in range(4):
        yyw = lines[i] @ P
        yyw_prime = lines[i] @ P_prime
        yyo = yyw @ wTo
        if not new_residual:
            if left:
                J[i] = -1 * (2 * yyo @ Qi @ wTo.T @ mu.circled_circ(yyw).T)
            else:
                J[i] = -2 * yyo @ Qi @ wTo.T @ cTw.T @ mu.circled_circ(yyw_prime).T
        else:
            uline_b = P.T @ lines[i]
            # NOTE the reference evaluates b from P = K*cTw here (ObjectResJacCam.cpp:446)
            b = uline_b[:3]
            b_norm = math.sqrt(float(b @ b))
            if left:
                p_ulineb_p_Cxi = wTo.T @ mu.circled_circ(yyw).T
            else:
                p_ulineb_p_Cxi = wTo.T @ cTw.T @ mu.circled_circ(yyw_prime).T
            dist = uline_b[3]
            sign = 1.0 if dist > 0 else -1.0
            sq = math.sqrt(float(b @ Qi[:3, :3] @ b))
            J[i] = -1 * (_new_bbox_chain(uline_b, b_norm, sq, sign, Qi) @ p_ulineb_p_Cxi)
    return J


def bbox_jac_object(cTw, wTo, shape, bbox, K, left, new_residual):
    """ObjectLM::ErrorBBoxQuadric::df per frame, ObjectLM.cpp:503-612 (4 x (9+3K))."""
    Qi = ellipse_from_shape(shape)
    P = cTw[:3, :]
    lines = poly2lineh(bbox2poly(bbox))
    J = np.zeros((4, 9 + 3 * K))
    shape = np.asarray(shape, dtype=float)
    for i in range(4):
        yyw = lines[i] @ P
        yyo = yyw @ wTo
        if not new_residual:
            if left:
                J[i, 0:6] = 2 * yyo @ Qi @ wTo.T @ mu.circled_circ(yyw).T
            else:
                J[i, 0:6] = 2 * yyo @ Qi @ mu.circled_circ(wTo.T @ yyw).T
            J[i, 6:9] = 2 * shape * (yyo[:3] ** 2)
        else:
            uline_b = P.T @ lines[i]
            b = uline_b[:3]
            b_norm = math.sqrt(float(b @ b))
            if left:
                p_ulineb_p_Oxi = wTo.T @ mu.circled_circ(yyw).T
            else:
                p_ulineb_p_Oxi = mu.circled_circ(wTo.T @ yyw).T
            dist = uline_b[3]
            sign = 1.0 if dist > 0 else -1.0
            sq = math.sqrt(float(b @ Qi[:3, :3] @ b))
            J[i, 0:6] = _new_bbox_chain(uline_b, b_norm, sq, sign, Qi) @ p_ulineb_p_Oxi
            J[i, 6:9] = (shape * (b * b)) / (b_norm * sq)
    return J


# ---------------------------------------------------------------- O3 / O4 stacking
def camera_lm(frames_wTc, wTo, shape, kps, zs_all, zb_all, left, new_residual,
              residual_weights=(1.0, 1.0)):
    """CameraLM::operator() and ::df, ObjectResJacCam.cpp:521-581 with huber = inf
    (identity, :606-643): rows = [all keypoint rows (frame-major); 4 bbox rows per
    frame]; Jacobian is rows x 6 (each row wrt ITS frame's camera pose).
    Returns (fvec, fjac, zs_num_wrt_timestamps, valid_camera_pose_mat 6xT)."""
    K = kps.shape[0]
    kps_h = np.hstack([kps, np.ones((K, 1))])
    f_kp, J_kp, f_bb, J_bb, zs_num = [], [], [], [], []
    for f, wTc in enumerate(frames_wTc):
        cTw = np.linalg.inv(wTc)
        f_kp.append(kp_residual(cTw, wTo, kps_h, zs_all[f]))
        J_kp.append(kp_jac_camera(cTw, wTo, kps_h, zs_all[f], left))
        f_bb.append(bbox_residual(cTw, wTo, shape, zb_all[f], new_residual))
        J_bb.append(bbox_jac_camera(cTw, wTo, shape, zb_all[f], left, new_residual))
        zs_num.append(len(valid_indices(zs_all[f])))
    fvec = np.concatenate([np.concatenate(f_kp) * residual_weights[0],
                           np.concatenate(f_bb) * residual_weights[1]])
    fjac = np.vstack([np.vstack(J_kp) * residual_weights[0],
                      np.vstack(J_bb) * residual_weights[1]])
    poses = np.stack([mu.se3_log(wTc) for wTc in frames_wTc], axis=1)
    return fvec, fjac, zs_num, poses


def object_lm_rows(frames_wTc, wTo, shape, kps, zs_all, zb_all, left, new_residual,
                   residual_weights=(1.0, 1.0), clone_pose_inverse=False):
    """Keypoint + bbox rows of ObjectLM::operator()/df (ObjectLM.cpp:761-816), i.e. what
    ObjectFeatureInitializer.cpp:424-432 keeps for the filter: (fvec, fjac rows x (9+3K)).
    clone_pose_inverse: cTw = [R^T | -R^T p] as ClonePose::getTransformGlobalToCam builds it
    (FeatureInitializer.h:51-81) instead of the matrix inverse -- they differ when R is only
    orthonormal to single precision, like the camera poses of the reference's one_car data."""
    K = kps.shape[0]
    kps_h = np.hstack([kps, np.ones((K, 1))])
    f_kp, J_kp, f_bb, J_bb = [], [], [], []
    for f, wTc in enumerate(frames_wTc):
        if clone_pose_inverse:
            cTw = np.eye(4)
            cTw[:3, :3] = wTc[:3, :3].T
            cTw[:3, 3] = -wTc[:3, :3].T @ wTc[:3, 3]
        else:
            cTw = np.linalg.inv(wTc)
        f_kp.append(kp_residual(cTw, wTo, kps_h, zs_all[f]))
        J_kp.append(kp_jac_object(cTw, wTo, kps_h, zs_all[f], left))
        f_bb.append(bbox_residual(cTw, wTo, shape, zb_all[f], new_residual))
        J_bb.append(bbox_jac_object(cTw, wTo, shape, zb_all[f], K, left, new_residual))
    fvec = np.concatenate([np.concatenate(f_kp) * residual_weights[0],
                           np.concatenate(f_bb) * residual_weights[1]])
    fjac = np.vstack([np.vstack(J_kp) * residual_weights[0],
                      np.vstack(J_bb) * residual_weights[1]])
    return fvec, fjac


# ---------------------------------------------------------------------------------------------------------------------
# Object pose initialisation (first step): ObjectFeatureInitializer::single_object_initialization
# without RANSAC (use_kabsch_with_ransac_flag = false, src/obj/ObjectFeatureInitializer.cpp:25-31, 99-111).
# Pinned by the reference's own known-answer tests src/tests/test_kabsch.cpp:10-87 (tests/test_oracle_cpu.py).
def find_transform(pts_in, pts_out):
    """findTransform, src/obj/ObjectFeatureInitializer.cpp:265-341: similarity transform (scale * R, t) that maps the
    3 x n points `pts_in` onto `pts_out` -- scale from the ratio of the polyline lengths, rotation by Kabsch (SVD of
    in * out^T with the reflection fix on the last singular direction).  Returns the 4 x 4 matrix."""
    a = np.array(pts_in, dtype=float)
    b = np.array(pts_out, dtype=float)
    assert a.shape == b.shape and a.shape[0] == 3
    n = a.shape[1]
    dist_in = sum(np.linalg.norm(a[:, c + 1] - a[:, c]) for c in range(n - 1))
    dist_out = sum(np.linalg.norm(b[:, c + 1] - b[:, c]) for c in range(n - 1))
    scale = dist_out / dist_in
    b = b / scale
    in_ctr = a.sum(axis=1) / n
    out_ctr = b.sum(axis=1) / n
    a = a - in_ctr[:, None]
    b = b - out_ctr[:, None]
    U, _, Vt = np.linalg.svd(a @ b.T)
    V = Vt.T
    d = 1.0 if np.linalg.det(V @ U.T) > 0 else -1.0
    R = V @ np.diag([1.0, 1.0, d]) @ U.T
    T = np.eye(4)
    T[:3, :3] = scale * R
    T[:3, 3] = scale * (out_ctr - R @ in_ctr)
    return T


def pose_se3_to_se2(T):
    """poseSE32SE2, include/orcvio/utils/se3_ops.hpp:272-300, literally: yaw = pi / atan2(r21, r11) (sic), z = 0."""
    out = np.eye(4)
    den = math.atan2(T[1, 0], T[0, 0])
    yaw = math.pi / den if den != 0.0 else float("inf")
    if not math.isfinite(yaw):
        yaw = 0.0
    out[0, 0] = math.cos(yaw)
    out[0, 1] = -math.sin(yaw)
    out[0, 3] = T[0, 3]
    out[1, 0] = math.sin(yaw)
    out[1, 1] = math.cos(yaw)
    out[1, 3] = T[1, 3]
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Object LM optimiser: the full ObjectLM functor (4 blocks), the keypoint triangulation of
# single_object_initialization and single_levenberg_marquardt.  Pinned by the reference's goldens
# test_error_deform_reg.h5 / test_error_mean_shape_reg.h5 (src/tests/test_object_lm.cpp:233-295) and by the assertions
# of its multi-frame tests on one_car/*.h5 (src/tests/test_object_lm_multiframe.cpp:61-125,
# test_object_init_multiframe.cpp:24-86) -- tests/test_oracle_cpu.py.
def deform_reg(kps, kps_mean, T):
    """ErrorDeformRegularization::operator()/df, src/obj/ObjectLM.cpp:652-718: (kps - mean), keypoint-major, repeated
    once per frame; Jacobian = identity on the keypoint's own columns."""
    K = kps.shape[0]
    f1 = (kps - kps_mean).reshape(-1)
    J1 = np.zeros((3 * K, 9 + 3 * K))
    J1[:, 9:] = np.eye(3 * K)
    return np.tile(f1, T), np.tile(J1, (T, 1))


def quadv_reg(shape, mean_shape, T, K):
    """ErrorQuadVRegularization::operator()/df, src/obj/ObjectLM.cpp:732-760."""
    J1 = np.zeros((3, 9 + 3 * K))
    J1[:, 6:9] = np.eye(3)
    return np.tile(np.asarray(shape, float) - mean_shape, T), np.tile(J1, (T, 1))


def object_lm_full(frames_wTc, wTo, shape, kps, zs_all, zb_all, left, new_residual, kps_mean, mean_shape, weights):
    """ObjectLM::operator() and ::df with all four functors (src/obj/ObjectLM.cpp:761-816), Huber epsilon = inf."""
    K = kps.shape[0]
    T = len(frames_wTc)
    f01, J01 = object_lm_rows(frames_wTc, wTo, shape, kps, zs_all, zb_all, left, new_residual,
                              residual_weights=(weights[0], weights[1]), clone_pose_inverse=True)
    f2, J2 = deform_reg(kps, kps_mean, T)
    f3, J3 = quadv_reg(shape, mean_shape, T, K)
    return (np.concatenate([f01, weights[2] * f2, weights[3] * f3]),
            np.vstack([J01, weights[2] * J2, weights[3] * J3]))


def object_state_plus(x, dx):
    """operator+(LMObjectState, Tangent), src/obj/ObjectLM.cpp:63-70, 211-227: the pose is ALWAYS retracted on the left
    (exp(dx) * wTo), whatever use_left_perturbation_flag says."""
    wTo, shape, kps = x
    return (mu.se3_exp(dx[:6]) @ wTo, shape + dx[6:9], kps + dx[9:].reshape(-1, 3))


def object_state_scaled_norm(diag, x):
    """LMObjectState::scaled_norm, include/orcvio/obj/ObjectLM.h:236-249: a SUM of the blocks' norms."""
    wTo, shape, kps = x
    s = float(np.linalg.norm(diag[:6] * mu.se3_log(wTo)))
    s += float(np.linalg.norm(diag[6:9] * shape))
    for k in range(kps.shape[0]):
        s += float(np.linalg.norm(kps[k] * diag[9 + 3 * k:12 + 3 * k]))
    return s


def triangulate_linear(uvs, wTcs):
    """single_triangulation_common, src/feat/FeatureInitializer.cpp:6-110: anchor = the LAST observing frame; rows
    B_perp(b_i) p = B_perp(b_i) p_CiinA with b_i the unit bearing rotated into the anchor frame; least squares."""
    R_GtoA = wTcs[-1][:3, :3].T
    p_AinG = wTcs[-1][:3, 3]
    A, b = [], []
    for uv, wTc in zip(uvs, wTcs):
        R_AtoCi = wTc[:3, :3].T @ R_GtoA.T
        p_CiinA = R_GtoA @ (wTc[:3, 3] - p_AinG)
        bi = R_AtoCi.T @ np.array([uv[0], uv[1], 1.0])
        bi = bi / np.linalg.norm(bi)
        Bp = np.array([[-bi[2], 0.0, bi[0]], [0.0, bi[2], -bi[1]]])
        A.append(Bp)
        b.append(Bp @ p_CiinA)
    p_f = np.linalg.lstsq(np.vstack(A), np.concatenate(b), rcond=None)[0]
    return R_GtoA.T @ p_f + p_AinG


def single_object_initialization(frames_wTc, zs_all, kps_mean, se2=True, min_obs=3):
    """ObjectFeatureInitializer::single_object_initialization without RANSAC (src/obj/ObjectFeatureInitializer.cpp:
    33-111): every keypoint seen in MORE than `min_obs` frames (ObjectFeature.cpp:86-127) is triangulated; with more
    than 3 such keypoints the pose is findTransform(mean, world) [+ poseSE32SE2].  Returns (ok, wTq, ids, points)."""
    ids, pts = [], []
    for k in range(kps_mean.shape[0]):
        fr = [f for f in range(len(frames_wTc)) if np.all(np.isfinite(zs_all[f][k]))]
        if len(fr) > min_obs:
            ids.append(k)
            pts.append(triangulate_linear([zs_all[f][k] for f in fr], [frames_wTc[f] for f in fr]))
    if len(ids) <= 3:
        return False, np.eye(4), ids, np.array(pts)
    T = find_transform(kps_mean[ids].T, np.array(pts).T)
    if se2:
        T = pose_se3_to_se2(T)
    return True, T, ids, np.array(pts)


def single_levenberg_marquardt(frames_wTc, zs_all, zb_all, wTo0, kps_mean, mean_shape, weights, left, new_residual):
    """ObjectFeatureInitializer::single_levenberg_marquardt, src/obj/ObjectFeatureInitializer.cpp:346-440: factor 10,
    start = (initial pose, mean shape, mean keypoints).  Returns the oracle.lm result dict (x = (wTo, shape, kps))
    plus `success` (lm.info() == Success: every status but ImproperInputParameters and TooManyFunctionEvaluation,
    LMonestep.h:90, 165-198)."""
    from . import lm

    def fun(x):
        return object_lm_full(frames_wTc, x[0], x[1], x[2], zs_all, zb_all, left, new_residual, kps_mean, mean_shape,
                              weights)[0]

    def jac(x):
        return object_lm_full(frames_wTc, x[0], x[1], x[2], zs_all, zb_all, left, new_residual, kps_mean, mean_shape,
                              weights)[1]

    x0 = (np.array(wTo0, float), np.array(mean_shape, float).ravel(), np.array(kps_mean, float))
    res = lm.minimize(fun, jac, x0, plus=object_state_plus, scaled_norm=object_state_scaled_norm, factor=10.0)
    res["success"] = res["status"] not in (lm.IMPROPER, lm.TOO_MANY_FEV)
    return res


def keypoints_to_global(kps, wTo):
    """transform_mean_keypoints_to_global, src/obj/ObjectState.cpp:15-40."""
    return kps @ wTo[:3, :3].T + wTo[:3, 3]


def displacement(T1, T2):
    """include/orcvio/utils/se3_ops.hpp:485-497."""
    return (3.0 - np.trace(T1[:3, :3].T @ T2[:3, :3])) / 2.0, float(np.linalg.norm(T1[:3, 3] - T2[:3, 3]))
