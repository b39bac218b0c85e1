"""CPU oracle -- frozen-frame "stack -> compress -> update" (TEST INFRASTRUCTURE, not product code).

Builds an OracleVIO whose window / covariance equal a synthetic frame snapshot and runs the
reference-shaped chain of removeLostFeatures (src/orcvio.cpp:2498-2560) followed by
measurementUpdate_hybrid (:1766-1950): per feature checkMotion + triangulate_position
(feature.hpp:353-396, 583-719), featureJacobian_msckf (:1171-1226), gatingTestFeature
(:1953-1976), dense stacking, Householder-QR compression when rows > cols (:2532-2552), EKF
update.  Dense like the reference.  Used by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py only.

parity: inherits oracle/filter.py's status (unpinned by the reference's own tests for these
functions; pinned by agreement between this oracle, the C++ restatement oracle/cpu_ref.cpp
and the CUDA path).
"""
from types import SimpleNamespace

import numpy as np

from . import feature as ofeat
from . import mathutils as mu
from .filter import OracleVIO, Feature, Clone

FL_LARVIO, FL_LEFT, FL_DISCARD = 1, 2, 4


def oracle_from_snapshot(snap, flags, sigma2, chi2_p=0.95, tri=None):
    """An OracleVIO whose window / covariance equal the snapshot (no yaml involved)."""
    vio = OracleVIO.__new__(OracleVIO)
    N = int(snap["n_clones"])
    p = SimpleNamespace(use_larvio_flag=int(bool(flags & FL_LARVIO)),
                        use_left_perturbation_flag=int(bool(flags & FL_LEFT)),
                        discard_large_update_flag=int(bool(flags & FL_DISCARD)),
                        feature_observation_noise=sigma2, estimate_td=False,
                        chi_square_threshold_feat=chi2_p, use_object_residual_update_cam_pose_flag=1)
    vio.p = p
    vio.LEG_DIM = 22
    vio.log = []
    vio.td = 0.0
    vio.state_cov = np.array(snap["P"], dtype=float).copy()
    vio.chi_table_feat = mu.chi2_table(chi2_p)
    vio.opt = ofeat.default_opt_config()
    if tri:
        for k, v in tri.items():
            setattr(vio.opt, k, v)
    R_b2c = np.array(snap["R_b2c"], dtype=float).reshape(3, 3)
    t_c_b = np.array(snap["t_c_b"], dtype=float).reshape(3)
    vio.clones = {}
    for c in range(N):
        cl = Clone(c)
        cl.orientation = np.array(snap["clone_R"][c], dtype=float).reshape(3, 3)
        cl.position = np.array(snap["clone_p"][c], dtype=float)
        cl.R_imu_cam0 = R_b2c
        cl.t_cam0_imu = t_c_b
        # same association order as the product's ob::m3_mulT / m3_vec so that the camera
        # poses fed to the bit-exact triangulation comparison are identical doubles
        R, B = cl.orientation, R_b2c
        cl.orientation_cam = np.array([[(R[i, 0] * B[j, 0] + R[i, 1] * B[j, 1]) + R[i, 2] * B[j, 2]
                                        for j in range(3)] for i in range(3)])
        cl.position_cam = cl.position + np.array(
            [(R[i, 0] * t_c_b[0] + R[i, 1] * t_c_b[1]) + R[i, 2] * t_c_b[2] for i in range(3)])
        vio.clones[c] = cl
    s = SimpleNamespace(id=N - 1, time=0.0, dt=0.0,
                        orientation=vio.clones[N - 1].orientation.copy(),
                        position=vio.clones[N - 1].position.copy(), velocity=np.zeros(3),
                        gyro_bias=np.zeros(3), acc_bias=np.zeros(3), R_imu_cam0=R_b2c.copy(),
                        t_cam0_imu=t_c_b.copy())
    vio.imu_state = s
    vio.map_server = {}
    vio.feature_states = []
    vio.grid_map = {}
    fo = snap["feat_off"]
    for f in range(len(fo) - 1):
        ft = Feature(f)
        for k in range(fo[f], fo[f + 1]):
            ft.observations[int(snap["obs_clone"][k])] = np.array(snap["obs_z"][k], dtype=float)
        vio.map_server[f] = ft
    return vio


def oracle_snapshot_update(snap, flags, sigma2, chi2_p=0.95, tri=None):
    """Reference-shaped 'stack -> compress -> update' on a frozen window, dense like the
    reference (removeLostFeatures :2498-2560 + measurementUpdate_hybrid)."""
    vio = oracle_from_snapshot(snap, flags, sigma2, chi2_p, tri)
    nf = len(vio.map_server)
    status = np.zeros(nf, dtype=np.int32)
    gamma = np.full(nf, -1.0)
    positions = np.zeros((nf, 3))
    iters = np.zeros((nf, 2), dtype=np.int32)
    cols = 22 + 6 * len(vio.clones)
    blocks = []
    for f in range(nf):
        ft = vio.map_server[f]
        ok = vio.checkMotion(ft, False) and vio._initialize(ft, None)
        if ft.tri_log is not None:
            iters[f] = (ft.tri_log.n_outer, ft.tri_log.n_inner_total)
        if not ok:
            continue
        status[f] |= 1
        positions[f] = ft.position
        sids = ft.obs_ids()
        H_xj, r_j = vio.featureJacobian_msckf(ft, sids)
        g = {}
        if vio.gatingTestFeature(H_xj, r_j, 2 * len(sids) - 3, g):
            status[f] |= 2
            blocks.append((H_xj[:, :cols], r_j))
        gamma[f] = g["gamma"]
    out = dict(status=status, gamma=gamma, positions=positions, iters=iters, vio=vio)
    if blocks:
        H = np.vstack([b[0] for b in blocks])
        r = np.concatenate([b[1] for b in blocks])
        out["H"] = H
        out["r"] = r
        if H.shape[0] > H.shape[1]:
            H, r = vio._compress(H, r, cols)
        Dn = vio.state_cov.shape[1]
        Hf = np.zeros((H.shape[0], Dn))
        Hf[:, :H.shape[1]] = H
        vio.measurementUpdate_hybrid(np.zeros((0, Dn)), np.zeros(0), np.zeros((0, Dn)), np.zeros(0), Hf, r)
        out["delta_x"] = vio.log[-1]["delta_x"]
        out["applied"] = vio.log[-1]["applied"]
    out["P"] = vio.state_cov
    return out
