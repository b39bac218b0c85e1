import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orcvio_b200 import api, synth
print('fp64 peak (DFMA, DMMA) TFLOP/s:', api.fp64_peak())
print('latency probe (cycles):', api.latency_probe())
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for (N, F, L, full) in [(20, 300, 6, False), (30, 1000, 6, False), (30, 2000, 6, False), (30, 4096, 6, False), (30, 256, 6, True)]:
    snap = synth.stress_snapshot(N, F, L, seed=1, full_tracks=full)
    api.snapshot_update(snap, flags=flags, noise_var=1.6e-5, cost_threshold=1e-3, init_final_dist_threshold=100.0, repeat=2)
    t0 = time.time()
    out = api.snapshot_update(snap, flags=flags, noise_var=1.6e-5, cost_threshold=1e-3, init_final_dist_threshold=100.0, repeat=20)
    wall = (time.time() - t0) / 20
    t = out['timings_us']
    print(f"N={N} F={F} full={full} pass={(out['status']&2).sum()//2}: tri {t[0]:.1f} jac {t[1]:.1f} qr_tiles {t[2]:.1f} qr_chain {t[3]:.1f} update {t[4]:.1f} total {t[5]:.1f} us (wall/rep {wall*1e6:.0f} us)")
