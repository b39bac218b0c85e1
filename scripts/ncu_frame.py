"""Runs the stress frame (BASELINE configs[3], case 4a) a few times on the resident frame handle:
the target of the `ncu --set full` captures summarised under profiles/."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orcvio_b200 import api, synth

n_feat = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
snap = synth.stress_snapshot(30, n_feat, 6, seed=0)
fr = api.Frame(30, 0, 1.6e-5, 0.95, -1.0, 1e-3, 100.0)
fr.load(snap)
for _ in range(reps):
    fr.run(1)
print("done", fr.kernel_launches())
