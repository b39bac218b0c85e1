"""Per-frame deviations GPU vs oracle in the hybrid mode (diagnostic; the assertions live in tests/test_gpu_hybrid_filter.py)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
from orcvio_b200 import api, synth
import helpers as H
from test_gpu_filter import _feed, _compare_decisions, _compare_state, _sync_oracle_from_gpu
from test_gpu_hybrid_filter import _sync_features, _compare_hybrid, CASES

case = int(sys.argv[1]) if len(sys.argv) > 1 else 0
config, overrides, n_frames, feats, n_landmarks, spec_kw = CASES[case]
seq = synth.make_sequence(synth.SynthSpec(config=config, seed=0, n_frames=n_frames, feats_per_frame=feats,
                                          overrides=overrides, n_landmarks=n_landmarks, **spec_kw))
vio = api.OrcVIO(H.write_cfg(seq["cfg"]))
assert vio.initialize()
it = H.run_oracle_sequence(seq)
state = dict(k=0)
counts = dict(ekf=0, ekf_rej=0, new=0, new_rej=0, lost=0, reanchor=0)
for fi in range(n_frames):
    _feed(vio, seq, fi, state)
    ref = next(it)
    msgs = []
    for name, fn in (("dec", lambda: _compare_decisions(fi, vio, ref)), ("hyb", lambda: _compare_hybrid(fi, vio, ref, counts)),
                     ("state", lambda: _compare_state(fi, vio, ref))):
        try:
            fn()
        except AssertionError as e:
            msgs.append(name + ": " + str(e).strip().splitlines()[0][:150])
        except Exception as e:
            msgs.append(name + ": EXC " + repr(e)[:150])
    P = vio.cov()
    dP = np.abs(P - ref.state_cov).max() / np.abs(ref.state_cov).max() if P.shape == ref.state_cov.shape else -1
    ids, anc, rho, oa, xyz = vio.feature_states()
    drho = max([abs(rho[k] - ref.map_server[int(f)].invDepth) / abs(rho[k]) for k, f in enumerate(ids)
                if int(f) in ref.map_server] + [0])
    if drho > 1e-9:
        base = 22 + 6 * len(ref.clones)
        hl = vio.hybrid_log()
        for k, f in enumerate(ids):
            ft = ref.map_server[int(f)]
            d = abs(rho[k] - ft.invDepth)
            if d > 1e-9 * abs(rho[k]):
                print(f"   feature {int(f)}: rho gpu {rho[k]:.12e} ref {ft.invDepth:.12e} abs {d:.2e} sigma {np.sqrt(P[base + k, base + k]):.3e} "
                      f"sigma_ref {np.sqrt(ref.state_cov[base + k, base + k]):.3e} new {int(f) in hl['new']} reanch {int(f) in hl['reanchored']}"
                      f" dPcol {np.abs(P[:, base + k] - ref.state_cov[:, base + k]).max():.2e}")
    print(f"frame {fi}: E gpu {len(ids)} ref {len(ref.feature_states)} dP {dP:.2e} drho {drho:.2e} zupt {vio.frame_stats().zupt}", *msgs, flush=True)
    try:
        _sync_oracle_from_gpu(ref, vio)
        _sync_features(ref, vio)
    except AssertionError as e:
        print("SYNC FAILED:", str(e)[:200])
        break
print(counts)
