"""Diagnostics: stage times of stress case 4b (m = N = 30) and wall-clock of the C2 object leg.

    python scripts/probe_4b_c2.py 4b | c2      (run under `ncu --metrics gpu__time_duration.sum` for the launch list)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402
from orcvio_b200 import api, configs, synth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "4b"
if what == "4b":
    snap = synth.stress_snapshot(bench.N_CLONES, bench.N_FEATURES, bench.MAX_TRACK, seed=0, full_tracks=True)
    fr = api.Frame(bench.N_CLONES, 0, bench.NOISE_VAR, 0.95, -1.0, bench.TRI["cost_threshold"],
                   bench.TRI["init_final_dist_threshold"])
    out = fr.update(fr.prepare_inputs(snap))
    fr.load(snap)
    for _ in range(3):
        fr.run(1)
    us = fr.run(5)
    _, st = fr.run(5, stages=True)
    print(json.dumps(dict(case="4b", us_per_frame=us, stages={k: round(v, 1) for k, v in st.items()},
                          kernel_times={k: round(v, 1) for k, v in fr.kernel_times().items()})))
else:
    print(json.dumps(bench._object_leg(api, configs, synth)))
