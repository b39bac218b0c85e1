#!/usr/bin/env python
"""Long-sequence per-update parity soak: the GPU filter (through the C ABI) against the oracle over whole
synthetic sequences, teacher-forced frame by frame exactly like tests/test_gpu_filter.py (same pre-frame state on
both sides; the filter is chaotic, so a free-running comparison measures the oracle's own sensitivity).

    python scripts/soak_parity.py [--frames 300] > profiles/rN_soak_parity.json

Every frame: identical candidate sets, triangulation validity, gate decisions, pruned clones (asserted), state and
covariance within 1e-9 (asserted); the JSON records the largest deviations seen.  Uses oracle/ (the checker) --
this is test infrastructure, not a product path."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(config, overrides, n_frames, feats, n_landmarks, spec_kw):
    from orcvio_b200 import api, synth
    import helpers as H
    import test_gpu_filter as T
    seq = synth.make_sequence(synth.SynthSpec(config=config, seed=0, n_frames=n_frames, feats_per_frame=feats,
                                              overrides=overrides, n_landmarks=n_landmarks, **spec_kw))
    vio = api.OrcVIO(H.write_cfg(seq["cfg"]))
    assert vio.initialize()
    it = H.run_oracle_sequence(seq)
    # conditioning of every update, measured on the oracle itself: a twin oracle takes the same frame from the same state
    # with the prior covariance moved by one unit in the last place (relative 1.1e-16, symmetric)
    it_b = H.run_oracle_sequence(seq)
    rng = np.random.default_rng(0)
    sens = dict(p_abs=0.0, v_abs=0.0, P_rel=0.0)
    over = []
    st = dict(k=0)
    n_cand = n_pass = n_zupt = 0
    worst = dict(P_rel=0.0, p_abs=0.0, v_abs=0.0, R_abs=0.0)
    p_gpu = []
    for fi in range(n_frames):
        if fi >= int(os.environ.get("SOAK_STOP", "1000000")):
            break
        T._feed(vio, seq, fi, st)
        ref = next(it)
        refb = next(it_b)
        sv = float(np.abs(refb.imu_state.velocity - ref.imu_state.velocity).max())
        sp = float(np.abs(refb.imu_state.position - ref.imu_state.position).max())
        sens["v_abs"] = max(sens["v_abs"], sv)
        sens["p_abs"] = max(sens["p_abs"], sp)
        sens["P_rel"] = max(sens["P_rel"], float(np.abs(refb.state_cov - ref.state_cov).max() / np.abs(ref.state_cov).max()))
        c, p = T._compare_decisions(fi, vio, ref, gamma_rtol=float(os.environ.get("SOAK_GAMMA_RTOL", "1e-7")))
        n_cand += c
        n_pass += p
        if not os.environ.get("SOAK_NO_ASSERT"):
            T._compare_state(fi, vio, ref)
        s, rs = vio.state(), ref.imu_state
        P = vio.cov()
        worst["P_rel"] = max(worst["P_rel"], float(np.abs(P - ref.state_cov).max() / np.abs(ref.state_cov).max()))
        worst["p_abs"] = max(worst["p_abs"], float(np.abs(np.array(s.p) - rs.position).max()))
        dv = float(np.abs(np.array(s.v) - rs.velocity).max())
        if dv > worst["v_abs"]:
            worst["v_frame"] = fi
        worst["v_abs"] = max(worst["v_abs"], dv)
        if os.environ.get("SOAK_NO_ASSERT") and dv > 5e-10:
            print(f"frame {fi}: dv {dv:.2e} dp {float(np.abs(np.array(s.p) - rs.position).max()):.2e} clones {s.n_clones} "
                  f"stats {vio.frame_stats().n_gate_pass_lost}/{vio.frame_stats().n_gate_pass_prune}", file=sys.stderr)
        worst["R_abs"] = max(worst["R_abs"], float(np.abs(np.array(s.R).reshape(3, 3) - rs.orientation).max()))
        n_zupt += int(vio.frame_stats().zupt)
        if dv > 1e-9:
            over.append(dict(frame=fi, gpu_vs_oracle_v=dv, oracle_vs_ulp_perturbed_oracle_v=sv))
        T._sync_oracle_from_gpu(ref, vio)
        T._sync_oracle_from_gpu(refb, vio)
        E = rng.normal(0.0, 1.1e-16, refb.state_cov.shape)
        refb.state_cov = refb.state_cov * (1.0 + (E + E.T) / 2)
        p_gpu.append(np.array(s.p))
    n_frames = len(p_gpu)
    gt = np.array([g[1] for g in seq["gt"][:n_frames]])
    ate = H.ate_first_pose_aligned(np.array(p_gpu), gt)
    return dict(config=config, overrides=overrides, frames=n_frames, features_per_frame=feats,
                candidates=n_cand, gated_in=n_pass, zupt_frames=n_zupt, worst_per_update=worst,
                oracle_sensitivity_to_one_ulp_of_P=sens, updates_above_1e_9_in_v=over,
                ate_vs_synthetic_truth_m=ate, decisions="identical on every frame (asserted)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=300)
    args = ap.parse_args()
    cases = [
        ("euroc", dict(if_ZUPT_valid=0, max_features_in_one_grid=0), args.frames, 150, 12000, {}),
        ("unity", dict(if_ZUPT_valid=0), args.frames, 120, 12000, {}),
        ("kitti_odom", dict(max_features_in_one_grid=0), args.frames, 250, 40000, {}),
    ]
    if os.environ.get("SOAK_CASES"):
        cases = [cases[int(k)] for k in os.environ["SOAK_CASES"].split(",")]
    out = [run(*c) for c in cases]
    print(json.dumps(dict(soak="per-update parity, teacher forced (tests/test_gpu_filter.py method)", runs=out), indent=1))


if __name__ == "__main__":
    main()
