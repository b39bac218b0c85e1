"""Free-running GPU filter against the free-running oracle (no teacher forcing) on sequences of the shipped shapes:
ATE between the two, next to the oracle's own sensitivity (the same oracle with the initial position moved by 1e-13 m)
and the ATE of both against the synthetic truth.  Writes one JSON object (profiles/rN_free_running_ate.json)."""
import copy, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from orcvio_b200 import api, synth
import helpers as H

CASES = [("unity", dict(if_ZUPT_valid=0), 120, 120, 6000), ("euroc", dict(max_features_in_one_grid=0), 120, 120, 6000),
         ("euroc", {}, 120, 120, 6000), ("kitti_odom", {}, 100, 250, 20000)]
out = []
for config, ov, n_frames, feats, nlm in CASES:
    seq = synth.make_sequence(synth.SynthSpec(config=config, seed=2, n_frames=n_frames, feats_per_frame=feats, overrides=ov, n_landmarks=nlm))
    seq2 = copy.deepcopy(seq)
    seq2["cfg"]["initial_pos"] = [x + 1e-13 for x in seq["cfg"]["initial_pos"]]
    vio = api.OrcVIO(H.write_cfg(seq["cfg"]))
    assert vio.initialize()
    it_a, it_b = H.run_oracle_sequence(seq), H.run_oracle_sequence(seq2)
    k = 0
    pg, pa, pb = [], [], []
    for fi, (t_img, f) in enumerate(seq["frames"]):
        k1 = k
        while k1 < len(seq["imu"]) and seq["imu"][k1][0] <= t_img + 0.02:
            k1 += 1
        vio.push_imu(seq["imu"][k:k1])
        k = k1
        assert vio.processFeatures(t_img, f)
        a, b = next(it_a), next(it_b)
        pg.append(np.array(vio.state().p)); pa.append(a.imu_state.position.copy()); pb.append(b.imu_state.position.copy())
    pg, pa, pb = np.array(pg), np.array(pa), np.array(pb)
    gt = np.array([g[1] for g in seq["gt"]])
    d = np.linalg.norm(pg - pa, axis=1)
    rec = dict(config=config, overrides=ov, frames=n_frames, features_per_frame=feats,
               ate_gpu_vs_oracle_m=float(d.mean()), max_gpu_vs_oracle_m=float(d.max()),
               first_frames_gpu_vs_oracle_m=float(d[:10].max()),
               ate_oracle_vs_perturbed_oracle_m=float(np.linalg.norm(pb - pa, axis=1).mean()),
               ate_gpu_vs_truth_m=H.ate_first_pose_aligned(pg, gt), ate_oracle_vs_truth_m=H.ate_first_pose_aligned(pa, gt))
    print(rec, flush=True)
    out.append(rec)
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/free_running_ate.json", "w"), indent=1)
