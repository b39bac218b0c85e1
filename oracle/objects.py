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
                   residual_weights=(1.0, 1.0)):
    """Keypoint + bbox rows of ObjectLM::operator()/df (ObjectLM.cpp:761-816), i.e. what
    ObjectFeatureInitializer.cpp:424-432 keeps for the filter: (fvec, fjac rows x (9+3K))."""
    K = kps.shape[0]
    kps_h = np.hstack([kps, np.ones((K, 1))])
    f_kp, J_kp, f_bb, J_bb = [], [], [], []
    for f, wTc in enumerate(frames_wTc):
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
