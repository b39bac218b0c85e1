"""GPU: the multi-trajectory batch (orcvio_batch_*) advances B independent filters in lock-step
and gives, per filter, exactly what a single-filter handle gives on the same inputs."""
import numpy as np
import pytest

from orcvio_b200 import api, montecarlo as mc
import helpers as H

pytestmark = pytest.mark.gpu


def test_batch_equals_single_filters():
    ids = [0, 1, 2, 3, 4]
    ov = dict(if_ZUPT_valid=0)
    seqs = mc.make_sequences("unity", ids, 28, 60, ov, n_landmarks=3000)
    cfg = H.write_cfg(seqs[0]["cfg"])
    rec, b = mc.run_local(cfg, seqs, ids)
    assert np.all(rec[:, 7] == 1.0)
    assert b.kernel_launches() > 0 and b.feature_updates() > 100
    for i, s in enumerate(seqs):
        vio = api.OrcVIO(H.write_cfg(s["cfg"]))
        assert vio.initialize()
        k = 0
        for (t_img, feats) in s["frames"]:
            k1 = k
            while k1 < len(s["imu"]) and s["imu"][k1][0] <= t_img + 0.02:
                k1 += 1
            vio.push_imu(s["imu"][k:k1])
            k = k1
            assert vio.processFeatures(t_img, feats)
        st, sb = vio.state(), b.state(i)
        np.testing.assert_allclose(np.array(sb.p), np.array(st.p), rtol=0, atol=1e-12)
        np.testing.assert_allclose(np.array(sb.R), np.array(st.R), rtol=0, atol=1e-12)
        assert sb.n_clones == st.n_clones
        P1, P2 = vio.cov(), b.cov(i)
        assert np.abs(P1 - P2).max() <= 1e-12 * np.abs(P1).max()
        assert rec[i, 5] < 0.5                       # ATE vs synthetic ground truth [m]
    out = mc.gather_records(rec, len(ids), 0, 1)
    np.testing.assert_array_equal(out[:, 0], np.arange(len(ids)))


def _metrics_numpy(est, gt):
    """System::publishGroundtruth restated (ros_wrapper/src/orcvio/src/System.cpp:885-943), one trajectory."""
    from oracle import mathutils as mu

    def q2R(q):
        x, y, z, w = q
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                         [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                         [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])

    def T(p7):
        M = np.eye(4)
        M[:3, :3] = q2R(p7[3:])
        M[:3, 3] = p7[:3]
        return M

    A = T(gt[0]) @ np.linalg.inv(T(est[0]))
    e_ori, e_pos = [], []
    for k in range(len(est)):
        Tc = A @ T(est[k])
        e_pos.append(np.linalg.norm(Tc[:3, 3] - gt[k, :3]))
        Rd = Tc[:3, :3] @ q2R(gt[k, 3:]).T                     # q_corrected (x) q_gt^-1
        ang = np.arccos(np.clip((np.trace(Rd) - 1) / 2, -1, 1))
        e_ori.append(np.degrees(2 * np.sin(ang / 2)))          # 2 |vec(q)| of a rotation by `ang`
    e_pos = np.array(e_pos)
    return np.array([np.mean(e_ori), e_pos.mean(), np.sqrt((e_pos ** 2).mean()), e_pos[-1]])


def test_threaded_replay_equals_lock_step_loop_and_device_metrics():
    """orcvio_batch_replay from several host threads (one batch each) == the per-frame orcvio_batch_process loop, bit
    for bit; the on-device trajectory metrics == the reference logger's arithmetic."""
    ids = [0, 1, 2, 3, 4, 5, 6]
    seqs = mc.make_sequences("unity", ids, 24, 60, dict(if_ZUPT_valid=0), n_landmarks=3000, workers=2)
    cfg = H.write_cfg(seqs[0]["cfg"])
    rec0, b = mc.run_local(cfg, seqs, ids)
    rec1, info = mc.run_replay(cfg, seqs, ids, n_threads=3)
    assert info["n_batches"] == 3 and np.all(rec1[:, 7] == 1.0)
    np.testing.assert_array_equal(rec1[:, 2:5], rec0[:, 2:5])              # final positions: identical
    assert info["feature_updates"] == b.feature_updates()
    gt = mc.gt_poses(seqs, info["n_frames"])
    for i in range(len(ids)):
        ref = _metrics_numpy(info["poses"][i], gt[i])
        np.testing.assert_allclose(info["metrics"][i, 1:], ref[1:], rtol=1e-9, atol=1e-9)
        # (the restatement goes through arccos of the trace, which loses digits at small angles)
        np.testing.assert_allclose(info["metrics"][i, 0], ref[0], rtol=1e-5, atol=1e-7)
        # position part == the first-pose translation-only alignment only when the first poses agree in rotation
        assert info["metrics"][i, 1] < 0.5


def test_tcw_and_pose_log(tmp_path):
    seq = mc.make_sequences("unity", [0], 8, 40, dict(if_ZUPT_valid=0), n_landmarks=1500)[0]
    vio = api.OrcVIO(H.write_cfg(seq["cfg"]))
    assert vio.initialize()
    log = tmp_path / "state_est_geo_feat.txt"
    assert vio.set_pose_log(log) == 0
    k = 0
    for (t_img, feats) in seq["frames"]:
        k1 = k
        while k1 < len(seq["imu"]) and seq["imu"][k1][0] <= t_img + 0.02:
            k1 += 1
        vio.push_imu(seq["imu"][k:k1])
        k = k1
        assert vio.processFeatures(t_img, feats)
    st = vio.state()
    R, t = vio.getTcw()
    Rb = np.array(st.R).reshape(3, 3)
    T = np.array(seq["cfg"]["T_cam_imu"]).reshape(4, 4)
    R_b2c = T[:3, :3]
    t_c_b = -R_b2c.T @ T[:3, 3]
    np.testing.assert_allclose(R, Rb @ R_b2c.T, atol=1e-12)              # orientation_cam = R_b2w R_b2c^T (:950-955)
    np.testing.assert_allclose(t, np.array(st.p) + Rb @ t_c_b, atol=1e-12)
    vio.set_pose_log(None)
    rows = np.loadtxt(log)
    assert rows.shape == (len(seq["frames"]), 8)
    np.testing.assert_allclose(rows[-1, 1:4], np.array(st.p), rtol=1e-5, atol=1e-5)
    assert abs(np.linalg.norm(rows[-1, 4:]) - 1) < 1e-4 and np.all(np.diff(rows[:, 0]) > 0)


def test_results_do_not_depend_on_batch_composition_or_concurrency():
    """Hybrid mode (euroc.yaml as shipped): a trajectory gives the same bits whether its filter runs alone (one batch
    and one host thread per trajectory, all of them concurrently on the GPU), in pairs, or in one batch -- the split-K
    reductions have a fixed association and every cross-stream dependency is explicit (a delayed side stream once let
    k_chol_w_solve overwrite F_1 under k_imu_factor)."""
    ids = list(range(12))
    seqs = mc.make_sequences("euroc", ids, 75, 120, {}, n_landmarks=3000, workers=4)
    cfg = H.write_cfg(seqs[0]["cfg"])
    assert seqs[0]["cfg"]["max_features_in_one_grid"] == 1
    ref = None
    for n_threads in (12, 6, 1, 12):
        rec, info = mc.run_replay(cfg, seqs, ids, n_threads=n_threads)
        assert np.all(rec[:, 7] == 1.0)
        if ref is None:
            ref = info["poses"]
        else:
            np.testing.assert_array_equal(info["poses"], ref)


def test_kitti_relative_error_matches_the_reference_package_and_oracle():
    """orcvio_kitti_relative_error: the device kernel against the golden outputs of the reference's own
    vendored rpg_trajectory_evaluation and against the oracle, incl. a batch of trajectories, a length without samples
    and the "TransError(%)" summary."""
    from oracle import trajmetrics as tm
    from test_oracle_cpu import _kitti_cases
    mk, g = _kitti_cases()
    for name, n, step, seed, lengths in mk.CASES:
        gt, es = mk.synth(n, seed, step)
        gt2, es2 = mk.synth(n, seed + 10, step)
        L = list(lengths) + [1e6]                           # the last length has no sample
        out, summ = api.kitti_relative_error(np.stack([es, es2]), np.stack([gt, gt2]), L)
        for li, ln in enumerate(lengths):
            st = g[f"{name}_{ln}_stats"]
            assert int(out[0, li, 0]) == int(st[0])
            assert abs(out[0, li, 1] - st[1]) <= 1e-11 and abs(out[0, li, 3] - st[3]) <= 1e-11
            assert abs(out[0, li, 2] - st[2]) <= 1e-9
        assert np.all(out[:, -1, :] == 0.0)
        for t, (e_, g_) in enumerate(((es, gt), (es2, gt2))):
            rows, s_ref = tm.kitti_summary(e_, g_, L)
            np.testing.assert_allclose(out[t], rows, rtol=0, atol=1e-9)
            assert abs(summ[t] - s_ref) <= 1e-10
        # Umeyama alignment + absolute error (orcvio_trajectory_align_ate) against the package's outputs, and the relative
        # error with the sim3 scale
        for method in ("sim3", "se3"):
            al = api.trajectory_align_ate(np.stack([es, es2]), np.stack([gt, gt2]), method)
            ref = g[f"{name}_{method}"]
            assert abs(al["s"][0] - ref[0]) <= 1e-12 and np.abs(al["R"][0].ravel() - ref[1:10]).max() <= 1e-11
            assert np.abs(al["t"][0] - ref[10:13]).max() <= 1e-9
            assert abs(al["mean"][0] - ref[13]) <= 1e-11 and abs(al["rmse"][0] - ref[14]) <= 1e-11
            s2, R2, t2, mean2, rmse2 = tm.absolute_error(es2, gt2, method)
            assert abs(al["s"][1] - s2) <= 1e-12 and abs(al["mean"][1] - mean2) <= 1e-11
        al = api.trajectory_align_ate(np.stack([es, es2]), np.stack([gt, gt2]), "sim3")
        out_s, _ = api.kitti_relative_error(np.stack([es, es2]), np.stack([gt, gt2]), [lengths[0]], scale=al["s"])
        st = g[f"{name}_{lengths[0]}_sim3scale_stats"]
        assert int(out_s[0, 0, 0]) == int(st[0]) and abs(out_s[0, 0, 1] - st[1]) <= 1e-11
