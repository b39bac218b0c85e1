"""CPU oracle -- KITTI-style relative trajectory error (TEST INFRASTRUCTURE, not product code).

Restatement of what the reference's trajectory evaluation computes (python_scripts/trajectory_eval/traj_eval.py:61-90 ->
the vendored rpg_trajectory_evaluation: trajectory.py:341-377 compute_relative_error_at_subtraj_len, :309-339
write_kitti_errors_to_yaml; compute_trajectory_errors.py:10-67 compute_relative_error with T_cm = I, scale = 1;
trajectory_utils.py:11-46), SURVEY 8f rank 4.

parity: PINNED by outputs of the reference package itself, imported in the build container
(tests/golden/make_kitti_rel_golden.py -> tests/golden/kitti_rel_error.npz).
"""
import numpy as np


def quaternion_matrix(q):
    """transformations.py:1410-1429, quaternion = [x, y, z, w] (3 x 3 part)."""
    q = np.array(q[:4], dtype=np.float64)
    nq = float(q @ q)
    if nq < np.finfo(float).eps * 4.0:
        return np.eye(3)
    q = q * np.sqrt(2.0 / nq)
    q = np.outer(q, q)
    return np.array([[1.0 - q[1, 1] - q[2, 2], q[0, 1] - q[2, 3], q[0, 2] + q[1, 3]],
                     [q[0, 1] + q[2, 3], 1.0 - q[0, 0] - q[2, 2], q[1, 2] - q[0, 3]],
                     [q[0, 2] - q[1, 3], q[1, 2] + q[0, 3], 1.0 - q[0, 0] - q[1, 1]]])


def distance_from_start(p):
    """trajectory_utils.get_distance_from_start."""
    d = np.sqrt(np.sum(np.diff(p, axis=0) ** 2, axis=1))
    return np.concatenate(([0.0], np.cumsum(d)))


def comparison_indices(distances, dist, max_dist_diff):
    """trajectory_utils.compute_comparison_indices_length: for every start the FIRST index whose distance is closest
    to d + dist within max_dist_diff; starts without one are DROPPED from the list (and the caller then pairs entry k
    with start k: compute_trajectory_errors.py:30-31 -- kept as it is)."""
    n = len(distances)
    comps = []
    for idx in range(n):
        best, err = -1, max_dist_diff
        target = distances[idx] + dist
        for i in range(idx, n):
            e = abs(distances[i] - target)
            if e < err:
                best, err = i, e
            elif distances[i] - target > err:
                break
        if best != -1:
            comps.append(best)
    return comps


def align_umeyama(p_gt, p_es, known_scale):
    """align_trajectory.py:27-79 (yaw_only = False) as align_utils.alignSIM3 / alignSE3 call it: gt = s R est + t."""
    mu_m, mu_d = p_gt.mean(0), p_es.mean(0)
    M, Dd = p_gt - mu_m, p_es - mu_d
    n = p_gt.shape[0]
    Cm = 1.0 / n * (M.T @ Dd)
    sigma2 = 1.0 / n * float(np.sum(Dd * Dd))
    U, Dv, Vt = np.linalg.svd(Cm)
    V = Vt.T
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(V) < 0:
        S[2, 2] = -1
    R = U @ S @ V.T
    s = 1.0 if known_scale else 1.0 / sigma2 * float(np.trace(np.diag(Dv) @ S))
    return s, R, mu_m - s * (R @ mu_d)


def absolute_error(pose_es, pose_gt, method):
    """Trajectory.align_trajectory (trajectory.py:211-246, all frames) + compute_absolute_error's translation part
    (compute_trajectory_errors.py:70-72) + results_writer.compute_statistics: (s, R, t, mean, rmse)."""
    s, R, t = align_umeyama(pose_gt[:, :3], pose_es[:, :3], known_scale=(method == "se3"))
    e = np.sqrt(np.sum((pose_gt[:, :3] - (s * (pose_es[:, :3] @ R.T) + t)) ** 2, axis=1))
    return s, R, t, float(np.mean(e)), float(np.sqrt(e @ e / len(e)))


def relative_error(pose_es, pose_gt, dist, max_dist_diff=None, scale=1.0):
    """compute_relative_error with T_cm = I (scale: the alignment's, 1 unless sim3).  pose: (n, 7) = p, q xyzw.  Returns dict(trans, trans_perc,
    rot_deg, rot_deg_per_m) of per-sample arrays (empty when fewer than two samples)."""
    if max_dist_diff is None:
        max_dist_diff = 0.2 * dist
    comps = comparison_indices(distance_from_start(pose_gt[:, :3]), dist, max_dist_diff)
    out = dict(trans=[], trans_perc=[], rot_deg=[], rot_deg_per_m=[])
    if len(comps) < 2:
        return {k: np.array(v) for k, v in out.items()}

    def T(pose):
        M = np.eye(4)
        M[:3, :3] = quaternion_matrix(pose[3:7])
        M[:3, 3] = pose[:3]
        return M

    for idx, c in enumerate(comps):
        T_c1_c2 = np.linalg.inv(T(pose_es[idx])) @ T(pose_es[c])
        T_c1_c2[:3, 3] *= scale
        T_m1_m2 = np.linalg.inv(T(pose_gt[idx])) @ T(pose_gt[c])
        E = np.linalg.inv(T_m1_m2) @ T_c1_c2
        tn = float(np.linalg.norm(E[:3, 3]))            # (the rotation into the world frame does not change the norm)
        ang = float(np.degrees(np.arccos(min(1.0, max(-1.0, (np.trace(E[:3, :3]) - 1.0) / 2.0)))))
        out["trans"].append(tn)
        out["trans_perc"].append(tn / dist * 100.0)
        out["rot_deg"].append(ang)
        out["rot_deg_per_m"].append(ang / dist)
    return {k: np.array(v) for k, v in out.items()}


def kitti_summary(pose_es, pose_gt, lengths, scale=1.0):
    """Per length (samples, mean trans %, mean rot deg/m, mean trans m) and write_kitti_errors_to_yaml's
    "TransError(%)" = sum of the per-length means over the lengths with samples / (their number + 1e-5)."""
    rows = []
    tot, valid = 0.0, 0
    for L in lengths:
        r = relative_error(pose_es, pose_gt, float(L), scale=scale)
        n = len(r["trans"])
        if n:
            rows.append([n, float(np.mean(r["trans_perc"])), float(np.mean(r["rot_deg_per_m"])), float(np.mean(r["trans"]))])
            valid += 1
            tot += rows[-1][1]
        else:
            rows.append([0, 0.0, 0.0, 0.0])
    return np.array(rows), tot / (valid + 1e-5)
