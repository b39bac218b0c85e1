"""GPU parity of the filter in the hybrid MSCKF / EKF-SLAM mode (max_features_in_one_grid > 0: config/euroc.yaml and
config/kitti_odom.yaml as shipped) against the oracle, through the C ABI -- SURVEY 8a J5 / J6 / U2 / T2 / H4, 8f(1).

Per frame, both sides starting from the same pre-frame state (teacher forcing, see test_gpu_filter.py): identical
sets of lost / updated / new EKF-SLAM features, identical gate decisions (dof-2 gate of the features of the state, MSCKF
gate of the new ones), identical anchor changes (getNewAnchorId), identical feature_states order, and state, covariance
(with the feature block), inverse depths and world positions within 1e-9 relative."""
import numpy as np
import pytest

from orcvio_b200 import api, synth
import helpers as H
from test_gpu_filter import _feed, _compare_decisions, _compare_state, _sync_oracle_from_gpu

pytestmark = pytest.mark.gpu

CASES = [
    # EKF-SLAM features open 5 s after the initialisation (last_ZUPT_time, src/orcvio.cpp:2291): frame 50 on
    ("euroc", dict(), 85, 120, 6000, {}),
    ("kitti_odom", dict(), 85, 250, 20000, {}),
    # a stand-still interval after the features entered the state: the ZUPT frame drops them all (:3104-3115)
    ("euroc", dict(zupt_max_feature_dis=0.03), 100, 120, 6000, dict(stops=((16.5, 17.4),))),
]


def _sync_features(ref, vio):
    ids, anc, rho, oa, xyz = vio.feature_states()
    assert [int(i) for i in ids] == list(ref.feature_states), "feature_states differ"
    for k, fid in enumerate(ids):
        ft = ref.map_server[int(fid)]
        assert ft.in_state and ft.ekf_feature
        assert ft.id_anchor == int(anc[k]), f"anchor of {int(fid)} differs"
        ft.invDepth = float(rho[k])
        ft.obs_anchor = np.array([oa[k][0], oa[k][1], 1.0])
        ft.position = xyz[k].copy()


def _compare_hybrid(fi, vio, ref, counts):
    hl = vio.hybrid_log()
    logs = [l for l in ref.log if l.get("state_id") == ref.imu_state.id]
    lg = [l for l in logs if l["kind"] == "removeLostFeatures"]
    pl = [l for l in logs if l["kind"] == "prune"]
    if lg:
        lg = lg[0]
        assert sorted(hl["ekf_lost"]) == sorted(lg["ekf_lost"]), f"frame {fi}: lost EKF features differ"
        if not lg.get("zupt"):
            assert sorted(hl["ekf"].keys()) == sorted(lg["ekf"]), f"frame {fi}: features of the state differ"
            for fid, g in lg["gate_ekf"].items():
                knife = abs(g["gamma"] - g["chi2"]) <= 1e-9 * g["chi2"]
                assert knife or hl["ekf"][fid][0] == g["pass"], f"frame {fi}: EKF gate differs ({fid})"
                assert abs(hl["ekf"][fid][1] - g["gamma"]) <= 1e-7 * abs(g["gamma"]) + 1e-13
                counts["ekf"] += 1
                counts["ekf_rej"] += int(not g["pass"])
            assert sorted(hl["new"].keys()) == sorted(lg["gate_ekf_new"].keys()), f"frame {fi}: new EKF features differ"
            for fid, g in lg["gate_ekf_new"].items():
                knife = abs(g["gamma"] - g["chi2"]) <= 1e-9 * g["chi2"]
                assert knife or hl["new"][fid][0] == g["pass"], f"frame {fi}: new-feature gate differs ({fid})"
                assert abs(hl["new"][fid][1] - g["gamma"]) <= 1e-7 * abs(g["gamma"]) + 1e-13
                counts["new"] += 1
                counts["new_rej"] += int(not g["pass"])
        counts["lost"] += len(lg["ekf_lost"])
    if pl:
        assert hl["reanchored"] == {int(k): (int(a), int(b)) for k, (a, b) in pl[0]["reanchored"].items()}, \
            f"frame {fi}: anchor changes differ"
        counts["reanchor"] += len(pl[0]["reanchored"])
    # the features of the state after the frame
    ids, anc, rho, oa, xyz = vio.feature_states()
    assert [int(i) for i in ids] == list(ref.feature_states), f"frame {fi}: feature_states differ"
    for k, fid in enumerate(ids):
        ft = ref.map_server[int(fid)]
        assert ft.id_anchor == int(anc[k]), f"frame {fi}: anchor of {int(fid)} differs"
        assert abs(rho[k] - ft.invDepth) <= 1e-9 * abs(ft.invDepth) + 1e-12, f"frame {fi}: inverse depth of {int(fid)}"
        np.testing.assert_allclose(oa[k], ft.obs_anchor[:2], rtol=1e-9, atol=1e-12)
        # the world position is derived: p_w = R (obs_anchor, 1) / rho + t amplifies the 1e-9 of rho by depth / |p_w|
        np.testing.assert_allclose(xyz[k], ft.position, rtol=3e-8, atol=3e-8)


@pytest.mark.parametrize("config,overrides,n_frames,feats,n_landmarks,spec_kw", CASES)
def test_hybrid_sequence_parity_per_update(config, overrides, n_frames, feats, n_landmarks, spec_kw):
    seq = synth.make_sequence(synth.SynthSpec(config=config, seed=0, n_frames=n_frames, feats_per_frame=feats,
                                              overrides=overrides, n_landmarks=n_landmarks, **spec_kw))
    assert seq["cfg"]["max_features_in_one_grid"] == 1          # the shipped value
    vio = api.OrcVIO(H.write_cfg(seq["cfg"]))
    assert vio.initialize()
    oracle_iter = H.run_oracle_sequence(seq)
    state = dict(k=0)
    counts = dict(ekf=0, ekf_rej=0, new=0, new_rej=0, lost=0, reanchor=0)
    n_zupt = max_E = 0
    for fi in range(n_frames):
        _feed(vio, seq, fi, state)
        ref = next(oracle_iter)
        _compare_decisions(fi, vio, ref)
        _compare_hybrid(fi, vio, ref, counts)
        fs = vio.frame_stats()
        if seq["cfg"]["if_ZUPT_valid"]:
            assert bool(fs.zupt) == bool(ref.if_ZUPT), f"frame {fi}: ZUPT decision differs"
            n_zupt += int(fs.zupt)
        _compare_state(fi, vio, ref)
        max_E = max(max_E, len(ref.feature_states))
        _sync_oracle_from_gpu(ref, vio)
        _sync_features(ref, vio)
    print(counts, "max E", max_E, "zupt frames", n_zupt)
    assert counts["ekf"] > 100 and counts["new"] > 20 and counts["lost"] > 5 and counts["reanchor"] > 10
    assert max_E >= 15
    if spec_kw.get("stops"):
        assert n_zupt >= 3
