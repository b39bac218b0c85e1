"""Are the results of a trajectory independent of the batch it runs in?  Same sequences, different groupings."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orcvio_b200 import configs, montecarlo as mc
n_traj, n_frames = int(sys.argv[1]), int(sys.argv[2])
ov = eval(sys.argv[3]) if len(sys.argv) > 3 else {}
ids = list(range(n_traj))
seqs = mc.make_sequences("euroc", ids, n_frames, 150, ov, n_landmarks=3000, workers=os.cpu_count())
path = os.path.join(tempfile.mkdtemp(), "cfg.yaml")
configs.write_yaml(path, seqs[0]["cfg"])
base = None
for th in (1, 1, 3, 8):
    rec, info = mc.run_replay(path, seqs, ids, n_threads=th)
    p = info["poses"]
    if base is None:
        base = p
        continue
    d = np.abs(p - base).max(axis=(1, 2))
    first = [int(np.argmax(np.abs(p[i] - base[i]).max(axis=1) > 0)) if d[i] > 0 else -1 for i in range(n_traj)]
    print("   differing trajectories:", [i for i in range(n_traj) if d[i] > 0], "first frames", [f for f in first if f >= 0])
    print(f"threads {th}: trajectories that differ {int((d > 0).sum())} / {n_traj}, max |dp| {d.max():.3e}, first differing frame {sorted(set(first))[:6]}", flush=True)
