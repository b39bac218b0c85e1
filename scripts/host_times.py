"""Wall-clock split of the end-to-end frame call (orcvio_frame_update): prepare / launch / wait+fetch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from orcvio_b200 import api

for nf in (2000, 4096):
    snap = bench.make_frame(0, nf)
    fr = api.Frame(bench.N_CLONES, 0, bench.NOISE_VAR, 0.95, -1.0, bench.TRI["cost_threshold"],
                   bench.TRI["init_final_dist_threshold"])
    inp = fr.prepare_inputs(snap)
    out = fr.update(inp)
    acc = np.zeros(4)
    t0 = time.perf_counter()
    n = 100
    for _ in range(n):
        fr.update(inp, out)
        h = fr.host_times(); acc += np.array([h["prepare"], h["launch"], h["wait_fetch"], h["total"]])
    wall = (time.perf_counter() - t0) / n * 1e6
    print(nf, "features: wall %.1f us; prepare %.1f launch %.1f wait+fetch %.1f total %.1f" % ((wall,) + tuple(acc / n)))
