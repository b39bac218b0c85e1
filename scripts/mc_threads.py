"""Replay throughput of the multi-trajectory batch against the number of host threads / batches (diagnostic)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orcvio_b200 import api, configs, montecarlo as mc
n_traj = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 60
ov = eval(sys.argv[3]) if len(sys.argv) > 3 else {}
ids = list(range(n_traj))
t0 = time.perf_counter()
seqs = mc.make_sequences("euroc", ids, n_frames, 150, ov, n_landmarks=3000, workers=os.cpu_count())
print("gen", round(time.perf_counter() - t0, 1), "s", "cores", os.cpu_count(), flush=True)
path = os.path.join(tempfile.mkdtemp(), "cfg.yaml")
configs.write_yaml(path, seqs[0]["cfg"])
for th in ([int(x) for x in os.environ['MC_THREADS'].split(',')] if os.environ.get('MC_THREADS') else (1, 2, 4, 8, 16, 32)):
    if th > n_traj:
        break
    rec, info = mc.run_replay(path, seqs, ids, n_threads=th)
    print(f"threads {th:3d}: {info['seconds']:.3f} s  {n_traj * n_frames / info['seconds']:9.0f} traj-frames/s  "
          f"thread-us per traj-frame {th * info['seconds'] / (n_traj * n_frames) * 1e6:7.1f}  launches {info['kernel_launches']}", flush=True)
