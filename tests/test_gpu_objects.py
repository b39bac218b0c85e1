"""GPU parity of stage 3 (object keypoint / bounding-box residual update) and of the stand-alone
stage 6 entry, through the C ABI.

  * orcvio_object_residuals (O1-O4) vs the oracle (oracle/objects.py, itself pinned to the
    reference's golden vectors) and directly vs the reference goldens (1e-6, the reference's bound);
  * orcvio_construct_object_jacobians (O5) vs the closed-form expectation of the reference's own
    test (src/tests/test_state_update.cpp:16-103) and vs the oracle with real camera Jacobians;
  * orcvio_remove_lost_objects (O6) vs the oracle: identical outcome, gamma 1e-8 rel, state and
    covariance within 1e-9 relative;
  * orcvio_propagate vs the oracle's processModel, 1e-12 relative.
"""
import os

import numpy as np
import pytest

from oracle import objects as obj
from oracle import mathutils as mu
from oracle.filter import OracleVIO
from orcvio_b200 import api, configs, synth
import helpers as H
from test_gpu_filter import _feed, _sync_oracle_from_gpu

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _look_at(cam_pos, target):
    z = target - cam_pos
    z = z / np.linalg.norm(z)
    x = np.cross(np.array([0.0, 0.0, 1.0]), z)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    T = np.eye(4)
    T[:3, :3] = np.column_stack([x, y, z])
    T[:3, 3] = cam_pos
    return T


def _object_scene(T=6, seed=0, drop=True):
    g = np.load(os.path.join(GOLD, "one_car.npz"))
    rng = np.random.default_rng(seed)
    kps = g["mean_shape"][0] + rng.normal(0, 0.02, (12, 3))
    shape = g["ellipsoid_shape"][0].ravel()
    wTo = mu.se3_exp(np.array([0.4, -0.3, 0.2, 0.05, -0.08, 0.6]))
    frames, zs, zb = [], [], []
    for f in range(T):
        ang = 0.25 * f
        cam = wTo[:3, 3] + np.array([9.0 * np.cos(ang), 9.0 * np.sin(ang), 1.5 + 0.1 * f])
        wTc = _look_at(cam, wTo[:3, 3] + rng.normal(0, 0.1, 3))
        cTw = np.linalg.inv(wTc)
        uv = obj.project_object_points(cTw[:3, :], wTo, np.hstack([kps, np.ones((12, 1))]))
        z = uv + rng.normal(0, 0.003, uv.shape)
        box = np.array([uv[:, 0].min(), uv[:, 1].min(), uv[:, 0].max(), uv[:, 1].max()]) + rng.normal(0, 0.004, 4)
        if drop:
            z[rng.choice(12, size=2 + f % 3, replace=False)] = np.nan
        frames.append(wTc)
        zs.append(z)
        zb.append(box)
    return np.array(frames), wTo, shape, kps, np.array(zs), np.array(zb)


@pytest.mark.parametrize("left", [True, False])
@pytest.mark.parametrize("new_residual", [False, True])
def test_object_rows_match_oracle(left, new_residual):
    frames, wTo, shape, kps, zs, zb = _object_scene()
    out = api.object_residuals(frames, wTo, shape, kps, zs, zb, left, new_residual)
    fvec, fjac, zs_num, poses = obj.camera_lm(frames, wTo, shape, kps, zs, zb, left, new_residual)
    fvec_o, fjac_o = obj.object_lm_rows(frames, wTo, shape, kps, zs, zb, left, new_residual)
    assert list(out["zs_num"]) == list(zs_num)
    assert out["fvec"].shape == fvec.shape
    np.testing.assert_allclose(out["fvec"], fvec, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(out["fvec"], fvec_o, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(out["fjac_cam"], fjac, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(out["fjac_obj"], fjac_o, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(out["cam_pose_se3"], poses, rtol=1e-10, atol=1e-12)


def test_object_rows_match_reference_goldens():
    gk = np.load(os.path.join(GOLD, "test_error_feature_quadric.npz"))
    gb = np.load(os.path.join(GOLD, "test_error_bbox_quadric.npz"))
    # keypoint residual + object Jacobian (reference src/tests/test_object_lm.cpp:90-152)
    out = api.object_residuals(np.linalg.inv(gk["S"])[None], gk["T"], gb["v"], gk["M"][:, :3], gk["zs"][None],
                               gb["zb"].reshape(1, 4), left=True, new_residual=False)
    assert np.abs(out["fvec"][:24] - gk["error"].ravel()).max() < 1e-6
    assert np.abs(out["fjac_obj"][:24] - gk["jacobian"]).max() < 1e-6
    # bbox residual + object Jacobian (reference src/tests/test_object_lm.cpp:154-202)
    out = api.object_residuals(np.linalg.inv(gb["S"])[None], gb["T"], gb["v"], np.zeros((12, 3)), gb["zs"][None],
                               gb["zb"].reshape(1, 4), left=True, new_residual=False)
    r0 = 2 * int(out["zs_num"][0])
    assert np.abs(out["fvec"][r0:r0 + 4] - gb["error"].ravel()).max() < 1e-6
    assert np.abs(out["fjac_obj"][r0:r0 + 4] - gb["jacobian"]).max() < 1e-6


def test_construct_object_jacobians_closed_form():
    """The reference's own test (src/tests/test_state_update.cpp:16-103), through the C ABI."""
    vio = api.OrcVIO(H.write_cfg(configs.make("unity", if_ZUPT_valid=0)))
    assert vio.initialize()
    ts, zs_num = [0.0, 1.0], [1, 1]
    LEG, nclone, F = 15, 2, 2
    assert vio.setStateCov(LEG, nclone) == 0
    assert vio.setWinPoseTimestamps(ts) == 0
    assert vio.fixDcamposeDimuposeToI() == 0
    rng = np.random.default_rng(5)
    rows = F * 2 + F * 4
    r, Hf, Jc = rng.uniform(-1, 1, rows), rng.uniform(-1, 1, (rows, 45)), rng.uniform(-1, 1, (rows, 6))
    flag, Hx, Hf_o, r_o = vio.constructObjectResidualJacobians(Jc, ts, Hf, r, zs_num, np.zeros((6, 2)))
    r_t, Hf_t, Hx_t = np.zeros(rows), np.zeros((rows, 45)), np.zeros((rows, LEG + 6 * nclone))
    for i in range(rows):
        if i < F * 2:
            nr, nc = (i // 2) * 6 + (i % 2), (i // 2) * 6 + LEG
        else:
            j = i - F * 2
            nr, nc = (j // 4) * 6 + (j % 4) + 2, (j // 4) * 6 + LEG
        r_t[nr] = r[i]
        Hf_t[nr] = Hf[i]
        Hx_t[nr, nc:nc + 6] = Jc[i]
    assert flag
    assert Hx.shape == Hx_t.shape
    np.testing.assert_array_equal(r_o, r_t)
    np.testing.assert_array_equal(Hf_o, Hf_t)
    np.testing.assert_array_equal(Hx, Hx_t)


def _running_filter(n_frames=26):
    seq = synth.make_sequence(synth.SynthSpec(config="unity", seed=4, n_frames=n_frames, feats_per_frame=100,
                                              overrides=dict(if_ZUPT_valid=0)))
    vio = api.OrcVIO(H.write_cfg(seq["cfg"]))
    assert vio.initialize()
    it = H.run_oracle_sequence(seq)
    state = dict(k=0)
    ref = None
    for fi in range(n_frames):
        _feed(vio, seq, fi, state)
        ref = next(it)
        _sync_oracle_from_gpu(ref, vio)
    return vio, ref


@pytest.mark.parametrize("scale,expect", [(1.0, "updated"), (60.0, "gate_fail")])
def test_object_update_matches_oracle(scale, expect):
    """constructObjectResidualJacobians + removeLostObjects on a running filter: an object seen from
    five clones of the window (one timestamp outside it), rows from the stage-3 kernel."""
    vio, ref = _running_filter()
    poses, ids, times = vio.window()
    N = len(ids)
    assert N >= 10
    Rbc, tcb = ref.imu_state.R_imu_cam0, ref.imu_state.t_cam0_imu
    sel = [2, 4, 5, 7, N - 2]
    frames, ts = [], []
    for c in sel:
        R, p = poses[c][:9].reshape(3, 3), poses[c][9:]
        wTc = np.eye(4)
        wTc[:3, :3] = R @ Rbc.T
        wTc[:3, 3] = p + R @ tcb
        frames.append(wTc)
        ts.append(float(times[c]))
    frames.append(frames[-1].copy())
    ts.append(-5.0)                                   # not in the window: rows must be dropped
    frames = np.array(frames)
    g = np.load(os.path.join(GOLD, "one_car.npz"))
    rng = np.random.default_rng(3)
    kps = g["mean_shape"][0]
    shape = g["ellipsoid_shape"][0].ravel()
    centre = frames[2][:3, 3] + frames[2][:3, :3] @ np.array([0.3, 0.1, 9.0])
    wTo = np.eye(4)
    wTo[:3, :3] = mu.so3_exp(np.array([0.1, -0.2, 0.8]))
    wTo[:3, 3] = centre
    zs, zb = [], []
    for wTc in frames:
        uv = obj.project_object_points(np.linalg.inv(wTc)[:3, :], wTo, np.hstack([kps, np.ones((12, 1))]))
        zs.append(uv + scale * rng.normal(0, 0.004, uv.shape))
        zb.append(np.array([uv[:, 0].min(), uv[:, 1].min(), uv[:, 0].max(), uv[:, 1].max()]) +
                  scale * rng.normal(0, 0.004, 4))
    zs, zb = np.array(zs), np.array(zb)
    zs[1, 3] = np.nan
    left = bool(ref.p.use_left_perturbation_flag)
    rows = api.object_residuals(frames, wTo, shape, kps, zs, zb, left=left, new_residual=True)
    flag, Hx, Hf, res = vio.constructObjectResidualJacobians(rows["fjac_cam"], ts, rows["fjac_obj"], rows["fvec"],
                                                             rows["zs_num"], rows["cam_pose_se3"])
    flag_o, Hx_o, Hf_o, res_o = ref.constructObjectResidualJacobians(rows["fjac_cam"], ts, rows["fjac_obj"],
                                                                      rows["fvec"], list(rows["zs_num"]),
                                                                      rows["cam_pose_se3"])
    assert flag and flag_o
    assert Hx.shape == Hx_o.shape and Hx.shape[0] == sum(2 * int(k) + 4 for k in rows["zs_num"][:-1])
    np.testing.assert_allclose(Hx, Hx_o, rtol=1e-10, atol=1e-12)
    np.testing.assert_array_equal(Hf, Hf_o)
    np.testing.assert_array_equal(res, res_o)
    status, gamma = vio.removeLostObjects(Hx, Hf, res)
    outcome = ref.removeLostObjects(Hx_o, Hf_o, res_o)
    names = {0: "updated", 1: "empty", 2: "disabled", 3: "nullspace_fail", 4: "gate_fail", 5: "nan"}
    assert names[status] == outcome == expect
    glog = [l for l in ref.log if l["kind"].startswith("object_gate")][-1]
    assert abs(gamma - glog["gamma"]) <= 1e-8 * abs(glog["gamma"])
    st, rs = vio.state(), ref.imu_state
    np.testing.assert_allclose(np.array(st.p), rs.position, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(np.array(st.v), rs.velocity, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(np.array(st.R).reshape(3, 3), rs.orientation, rtol=0, atol=1e-9)
    P = vio.cov()
    assert np.abs(P - ref.state_cov).max() <= 1e-9 * np.abs(ref.state_cov).max()
    poses2, ids2, _ = vio.window()
    for c, sid in enumerate(ids2):
        np.testing.assert_allclose(poses2[c][9:], ref.clones[int(sid)].position, rtol=1e-9, atol=1e-9)


def test_object_update_edge_cases():
    vio, _ = _running_filter(12)
    D = vio.cov().shape[0]
    assert vio.removeLostObjects(np.zeros((0, D)), np.zeros((0, 45)), np.zeros(0))[0] == 1       # empty
    assert vio.removeLostObjects(np.zeros((10, D)), np.zeros((10, 45)), np.zeros(10))[0] == 3    # rows <= cols


@pytest.mark.parametrize("config", ["unity", "euroc"])
def test_propagate_matches_oracle(config):
    ov = dict(if_ZUPT_valid=0, max_features_in_one_grid=0)
    seq = synth.make_sequence(synth.SynthSpec(config=config, seed=1, n_frames=3, feats_per_frame=20, overrides=ov))
    ref = OracleVIO(H.write_cfg(seq["cfg"]))
    assert ref.initialize()
    rng = np.random.default_rng(0)
    N = 4
    D = 22 + 6 * N
    A = rng.normal(0, 0.05, (D, D))
    P = A @ A.T + np.diag(np.r_[np.full(3, 4e-4), np.full(3, 0.2), np.full(3, 0.5), np.full(3, 4e-4), np.full(3, 0.01),
                                np.zeros(7), np.full(6 * N, 0.01)])
    P[15:22, :] = 0
    P[:, 15:22] = 0
    s = ref.imu_state
    s.orientation = mu.so3_exp(np.array([0.3, -0.2, 0.5]))
    s.velocity = np.array([0.4, -0.2, 0.1])
    s.position = np.array([1.0, 2.0, 0.5])
    s.gyro_bias = np.array([0.002, -0.001, 0.0015])
    s.acc_bias = np.array([0.02, 0.01, -0.015])
    s.time = float(seq["imu"][4][0])
    ref.state_cov = P.copy()
    ref.m_gyro_old = seq["imu"][4][1:4].copy()
    ref.m_acc_old = seq["imu"][4][4:7].copy()
    ref.imu_state_old = ref._copy_imu(s)
    samples = seq["imu"][5:30]
    flags = (1 if ref.p.use_larvio_flag else 0) | (2 if ref.p.use_left_perturbation_flag else 0)
    noise4 = [ref.p.imu_gyro_noise, ref.p.imu_acc_noise, ref.p.imu_gyro_bias_noise, ref.p.imu_acc_bias_noise]
    R1, v1, p1, t1, P1 = api.propagate(s.orientation, s.velocity, s.position, s.time, s.gyro_bias, s.acc_bias,
                                       ref.m_gyro_old, ref.m_acc_old, samples, P, flags, noise4)
    buf = [(r[0], r[1:4].copy(), r[4:7].copy()) for r in samples]
    ref.batchImuProcessing(samples[-1][0], buf)
    assert t1 == ref.imu_state.time
    np.testing.assert_allclose(R1, ref.imu_state.orientation, rtol=0, atol=1e-12)
    np.testing.assert_allclose(v1, ref.imu_state.velocity, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(p1, ref.imu_state.position, rtol=1e-12, atol=1e-12)
    assert np.abs(P1 - ref.state_cov).max() <= 1e-12 * np.abs(ref.state_cov).max()


def test_kabsch_init_matches_reference_known_answers_and_oracle():
    """orcvio_object_kabsch_init: the reference's own known-answer tests (src/tests/test_kabsch.cpp, 1e-13) through the
    C ABI, a batch of random objects against the oracle, the SE(2) projection, a degenerate object."""
    from test_oracle_cpu import _kabsch_kat
    rng = np.random.default_rng(5)
    means, worlds, want = [], [], []
    for planar in (False, True):
        pin, pout, sR, S = _kabsch_kat(planar)
        means.append(pin.T)
        worlds.append(pout.T)
        want.append((sR, S))
    g = np.load(os.path.join(GOLD, "one_car.npz"))
    kps = g["mean_shape"][0]
    for k in range(30):
        sel = np.sort(rng.choice(12, size=int(rng.integers(4, 13)), replace=False))
        R = mu.so3_exp(rng.normal(0, 1.0, 3))
        t = rng.normal(0, 5.0, 3)
        w = (R @ kps[sel].T).T + t + rng.normal(0, 0.02, (len(sel), 3))
        means.append(kps[sel])
        worlds.append(w)
    T, ok = api.object_kabsch_init(means, worlds)
    assert np.all(ok == 1)
    for i, (sR, S) in enumerate(want):
        assert np.abs(T[i, :3, :3] - sR).max() <= 1e-13 and np.abs(T[i, :3, 3] - S).max() <= 1e-13
    for i in range(2, len(means)):
        ref = obj.find_transform(means[i].T, worlds[i].T)
        np.testing.assert_allclose(T[i], ref, rtol=0, atol=1e-11)
    T2, ok2 = api.object_kabsch_init(means, worlds, se2=True)
    for i in range(len(means)):
        np.testing.assert_allclose(T2[i], obj.pose_se3_to_se2(T[i]), rtol=0, atol=1e-9)
    T3, ok3 = api.object_kabsch_init([kps[:1], kps[:5]], [kps[:1], np.tile(kps[:1], (5, 1))])
    assert list(ok3) == [0, 0]


# ------------------------------------------------------------------ object state optimiser
def _lm_scenes(n, seed=3):
    """Objects of one class seen over 5..14 frames with dropped keypoints; plus the reference's one_car sequence."""
    from test_oracle_cpu import _one_car
    d = _one_car()
    scenes = [(d["frames"], d["zs"], d["zb"])]
    for i in range(n):
        fr, wTo, shape, kps, zs, zb = _object_scene(T=5 + (3 * i) % 10, seed=seed + i, drop=True)
        scenes.append((fr, zs, zb))
    return d, scenes


@pytest.mark.parametrize("left", [True, False])
@pytest.mark.parametrize("new_residual", [False, True])
def test_object_lm_normal_equations_match_oracle(left, new_residual):
    """k_object_lm_eval: |f|, J^T f, J^T J of the four-block ObjectLM functor against the oracle's stacked Jacobian
    (1e-12 relative to the largest entry), ragged observations, weights all different."""
    d, scenes = _lm_scenes(6)
    rng = np.random.default_rng(11)
    w = [1.0, 3e-2, 0.7, 1.3]
    init = api.ObjectFeatureInitializer(d["mean_shape"], d["kps_mean"], w)
    states = []
    for (fr, zs, zb) in scenes:
        ok, T0, _, _ = obj.single_object_initialization(fr, zs, d["kps_mean"], se2=False)
        T0[:3, :3] /= np.cbrt(np.linalg.det(T0[:3, :3]))
        states.append((mu.se3_exp(rng.normal(0, 0.03, 6)) @ T0, d["mean_shape"] + rng.normal(0, 0.05, 3),
                       d["kps_mean"] + rng.normal(0, 0.02, (12, 3))))
    fn, g, A = init.lm_eval([s[0] for s in scenes], [s[1] for s in scenes], [s[2] for s in scenes], states, left,
                            new_residual)
    for i, ((fr, zs, zb), x) in enumerate(zip(scenes, states)):
        f, J = obj.object_lm_full(fr, x[0], x[1], x[2], zs, zb, left, new_residual, d["kps_mean"], d["mean_shape"], w)
        assert abs(fn[i] - np.linalg.norm(f)) <= 1e-12 * np.linalg.norm(f)
        JtJ, Jtf = J.T @ J, J.T @ f
        assert np.abs(A[i] - JtJ).max() <= 1e-12 * np.abs(JtJ).max(), i
        assert np.abs(g[i] - Jtf).max() <= 1e-12 * np.abs(Jtf).max(), i


def test_object_initialisation_matches_oracle():
    """orcvio_object_init (keypoint triangulation + Kabsch + poseSE32SE2) against the oracle on the reference's two
    sequences and on ragged synthetic objects, incl. one with too few triangulable keypoints."""
    from test_oracle_cpu import _one_car
    d, scenes = _lm_scenes(8)
    scenes.append((_one_car("one_car_no_zb")["frames"], _one_car("one_car_no_zb")["zs"], None))
    few = np.array(scenes[1][1], copy=True)
    few[:, 3:, :] = np.nan                       # only 3 keypoints left: no pose
    scenes.append((scenes[1][0], few, None))
    init = api.ObjectFeatureInitializer(d["mean_shape"], d["kps_mean"])
    for se2 in (True, False):
        init.estimate_SE2_pose_flag = se2
        ok, T, kw, kv = init.single_object_initialization([s[0] for s in scenes], [s[1] for s in scenes])
        for i, (fr, zs, _) in enumerate(scenes):
            ok_o, T_o, ids, pts = obj.single_object_initialization(fr, zs, d["kps_mean"], se2=se2)
            assert bool(ok[i]) == ok_o and list(np.nonzero(kv[i])[0]) == ids
            if len(ids):
                np.testing.assert_allclose(kw[i][ids], pts, rtol=0, atol=1e-9 * max(1.0, np.abs(pts).max()))
            np.testing.assert_allclose(T[i], T_o, rtol=0, atol=1e-8)
    assert list(ok[-1:]) == [0]


@pytest.mark.parametrize("left,new_residual", [(True, False), (False, False), (True, True)])
def test_object_lm_matches_oracle(left, new_residual):
    """orcvio_object_lm for a batch in lock-step against the oracle's restatement of the reference optimiser run one
    object at a time: same status, nfev, njev; optimum within 1e-7 (the two differ by a column-pivoted QR of J against
    a pivoted Cholesky of J^T J and stop on sqrt(eps) tests)."""
    d, scenes = _lm_scenes(5)
    w = [1.0, 3e-2, 1.0, 1.0]
    init = api.ObjectFeatureInitializer(d["mean_shape"], d["kps_mean"], w)
    ok, T0, _, _ = init.single_object_initialization([s[0] for s in scenes], [s[1] for s in scenes])
    assert np.all(ok == 1)
    out = init.single_levenberg_marquardt([s[0] for s in scenes], [s[1] for s in scenes], [s[2] for s in scenes], T0,
                                          left, new_residual)
    assert out["rounds"] == out["nfev"].max()
    for i, (fr, zs, zb) in enumerate(scenes):
        res = obj.single_levenberg_marquardt(fr, zs, zb, T0[i], d["kps_mean"], d["mean_shape"], w, left, new_residual)
        assert bool(out["success"][i]) == res["success"]
        assert (out["status"][i], out["nfev"][i], out["njev"][i]) == (res["status"], res["nfev"], res["njev"]), i
        assert abs(out["fnorm"][i] - res["fnorm"]) <= 1e-9 * max(1.0, res["fnorm"])
        np.testing.assert_allclose(out["wTo"][i], res["x"][0], rtol=0, atol=1e-7)
        np.testing.assert_allclose(out["shape"][i], res["x"][1], rtol=0, atol=1e-7)
        np.testing.assert_allclose(out["kps"][i], res["x"][2], rtol=0, atol=1e-7)
        np.testing.assert_allclose(out["kps_world"][i], obj.keypoints_to_global(res["x"][2], res["x"][0]), rtol=0, atol=1e-6)
    # the reference's own assertion on its sequence (test_object_lm_multiframe.cpp:115-123)
    dR, dt = obj.displacement(d["wTq"], out["wTo"][0])
    assert abs(dR) < 0.5 and dt < 0.05 * np.linalg.norm(d["wTq"][:3, 3])


@pytest.mark.parametrize("K", [4, 8])
def test_object_optimiser_with_smaller_keypoint_classes(K):
    """The reference's classes have 12, 4 and 8 keypoints (config/object_feat_*.yaml): the first K keypoints of the
    car shape as a class of its own -- initialisation, normal equations and the optimum against the oracle."""
    d, scenes = _lm_scenes(4)
    kps_mean = d["kps_mean"][:K]
    scenes = [(fr, zs[:, :K], zb) for (fr, zs, zb) in scenes]
    w = [1.0, 3e-2, 1.0, 1.0]
    init = api.ObjectFeatureInitializer(d["mean_shape"], kps_mean, w)
    ok, T0, kw, kv = init.single_object_initialization([s[0] for s in scenes], [s[1] for s in scenes])
    for i, (fr, zs, _) in enumerate(scenes):
        ok_o, T_o, ids, pts = obj.single_object_initialization(fr, zs, kps_mean)
        assert bool(ok[i]) == ok_o and list(np.nonzero(kv[i])[0]) == ids
        np.testing.assert_allclose(T0[i], T_o, rtol=0, atol=1e-8)
    sel = [i for i in range(len(scenes)) if ok[i]]
    assert len(sel) >= 2
    fl, zl, bl = [scenes[i][0] for i in sel], [scenes[i][1] for i in sel], [scenes[i][2] for i in sel]
    out = init.single_levenberg_marquardt(fl, zl, bl, T0[sel], True, False)
    n = 9 + 3 * K
    fn, g, A = init.lm_eval(fl, zl, bl, [(out["wTo"][j], out["shape"][j], out["kps"][j]) for j in range(len(sel))])
    for j, i in enumerate(sel):
        fr, zs, zb = scenes[i]
        res = obj.single_levenberg_marquardt(fr, zs, zb, T0[i], kps_mean, d["mean_shape"], w, True, False)
        assert (out["status"][j], out["nfev"][j], out["njev"][j]) == (res["status"], res["nfev"], res["njev"]), i
        np.testing.assert_allclose(out["wTo"][j], res["x"][0], rtol=0, atol=1e-7)
        np.testing.assert_allclose(out["kps"][j], res["x"][2], rtol=0, atol=1e-7)
        f, J = obj.object_lm_full(fr, out["wTo"][j], out["shape"][j], out["kps"][j], zs, zb, True, False, kps_mean,
                                  d["mean_shape"], w)
        assert A[j].shape == (n, n) and np.abs(A[j] - J.T @ J).max() <= 1e-12 * np.abs(J.T @ J).max()
