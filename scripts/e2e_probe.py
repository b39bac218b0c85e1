"""Wall-clock split of orcvio_frame_update (host work-list build / launches / wait+fetch) on the stress frame."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orcvio_b200 import api, synth
for F in (2000, 4096):
    snap = synth.stress_snapshot(30, F, 6, seed=0)
    fr = api.Frame(30, 0, 1.6e-5, 0.95, -1.0, 1e-3, 100.0)
    inp = fr.prepare_inputs(snap)
    out = fr.update(inp)
    acc = {}
    for _ in range(20):
        fr.update(inp, out)
        for k, v in fr.host_times().items():
            acc[k] = acc.get(k, 0.0) + v / 20
    print(F, {k: round(v, 1) for k, v in acc.items()})
