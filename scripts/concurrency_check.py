"""Does a frame update give the same bits when many handles run concurrently on one GPU (one host thread each)?"""
import os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from orcvio_b200 import api, synth
n_threads = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
n_clones, n_feat = 20, 150
snap = synth.stress_snapshot(n_clones, n_feat, 6, seed=0, config="euroc")

def mk():
    return api.Frame(n_clones, 0, bench.NOISE_VAR, 0.95, -1.0, bench.TRI["cost_threshold"], bench.TRI["init_final_dist_threshold"])

fr0 = mk()
ref = {k: v.copy() for k, v in fr0.update(fr0.prepare_inputs(snap)).items() if hasattr(v, "copy")}
bad = {}

def work(k):
    fr = mk()
    inp = fr.prepare_inputs(snap)
    out = None
    for r in range(reps):
        out = fr.update(inp, out)
        for key in ("P", "delta_x", "status", "gamma"):
            if not np.array_equal(out[key], ref[key]):
                d = float(np.abs(out[key].astype(float) - ref[key].astype(float)).max())
                bad.setdefault(key, []).append((k, r, d))
                break

ths = [threading.Thread(target=work, args=(k,)) for k in range(n_threads)]
for t in ths: t.start()
for t in ths: t.join()
print("threads", n_threads, "reps", reps, "mismatches:", {k: (len(v), v[:3]) for k, v in bad.items()})
