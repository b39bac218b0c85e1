"""Free-running GPU filter vs free-running oracle on euroc.yaml as shipped: per-frame distance and the first frame at
which a decision (candidate sets, gates, EKF feature sets) differs.  Meant to be run plain and under compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from orcvio_b200 import api, synth
import helpers as H
from test_gpu_filter import _feed, _compare_decisions
from test_gpu_hybrid_filter import _compare_hybrid
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seq = synth.make_sequence(synth.SynthSpec(config="euroc", seed=2, n_frames=n_frames, feats_per_frame=120, overrides={}, n_landmarks=6000))
vio = api.OrcVIO(H.write_cfg(seq["cfg"]))
assert vio.initialize()
it = H.run_oracle_sequence(seq)
state = dict(k=0)
counts = dict(ekf=0, ekf_rej=0, new=0, new_rej=0, lost=0, reanchor=0)
first = None
for fi in range(n_frames):
    _feed(vio, seq, fi, state)
    ref = next(it)
    d = np.linalg.norm(np.array(vio.state().p) - ref.imu_state.position)
    P = vio.cov()
    dP = np.abs(P - ref.state_cov).max() / np.abs(ref.state_cov).max() if P.shape == ref.state_cov.shape else -1
    msg = ""
    if first is None:
        for name, fn in (("dec", lambda: _compare_decisions(fi, vio, ref)), ("hyb", lambda: _compare_hybrid(fi, vio, ref, counts))):
            try:
                fn()
            except AssertionError as e:
                msg += f" {name}: " + str(e).strip().splitlines()[0][:160]
        if msg:
            first = fi
    print(f"frame {fi:3d} |dp| {d:.3e} dP {dP:.2e} E {len(vio.feature_states()[0])}/{len(ref.feature_states)}{msg}", flush=True)
