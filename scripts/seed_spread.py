"""Device time of the stress frame for different seeds (the frames of the ranks of a multi-GPU run): how much of the
max-over-ranks is the spread of the work itself (the Levenberg-Marquardt tail of the slowest feature)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from orcvio_b200 import api
buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for seed in range(8):
    snap = bench.make_frame(seed)
    fr = api.Frame(bench.N_CLONES, 0, bench.NOISE_VAR, 0.95, -1.0, bench.TRI["cost_threshold"], bench.TRI["init_final_dist_threshold"])
    fr.update(fr.prepare_inputs(snap))
    fr.load(snap)
    for _ in range(5):
        fr.run(1)
    tot = 0.0
    for _ in range(30):
        buf.zero_(); torch.cuda.synchronize()
        tot += fr.run(1)
    _, st = fr.run(5, stages=True)
    print(f"seed {seed}: {tot / 30:.1f} us per frame; tri {st['tri']:.1f} jac {st['jac_gate']:.1f}", flush=True)
    del fr
