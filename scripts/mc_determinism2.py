"""Hybrid mode: single-filter batches (the configuration the oracle parity tests cover) against larger batches."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orcvio_b200 import configs, montecarlo as mc
n_traj, n_frames = int(sys.argv[1]), int(sys.argv[2])
ids = list(range(n_traj))
seqs = mc.make_sequences("euroc", ids, n_frames, 150, {}, n_landmarks=3000, workers=os.cpu_count())
path = os.path.join(tempfile.mkdtemp(), "cfg.yaml")
configs.write_yaml(path, seqs[0]["cfg"])
runs = {}
for name, th in (("single", n_traj), ("single again", n_traj), ("pairs", n_traj // 2), ("one batch", 1), ("one batch again", 1)):
    rec, info = mc.run_replay(path, seqs, ids, n_threads=th)
    runs[name] = info["poses"]
base = runs["single"]
for name, p in runs.items():
    d = np.abs(p - base).max(axis=(1, 2))
    first = {i: int(np.argmax(np.abs(p[i] - base[i]).max(axis=1) > 0)) for i in range(n_traj) if d[i] > 0}
    print(f"{name:16s}: differ from single-filter runs: {len(first)} / {n_traj}  max |dp| {d.max():.3e}  first frames {first}", flush=True)
