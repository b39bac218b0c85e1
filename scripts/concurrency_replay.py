"""The same sequence through N single-filter batches replayed (orcvio_batch_replay) on N host threads."""
import os, sys, threading, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orcvio_b200 import api, configs, synth, montecarlo as mc
n_threads = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 100
inside = len(sys.argv) > 3 and sys.argv[3] == "inside"
seq = synth.make_sequence(synth.SynthSpec(config="euroc", seed=3, n_frames=n_frames, feats_per_frame=150,
                                          overrides=dict(max_features_in_one_grid=0), n_landmarks=3000))
path = os.path.join(tempfile.mkdtemp(), "cfg.yaml")
configs.write_yaml(path, seq["cfg"])
pack = mc.pack_replay([seq], n_frames)
out = [None] * n_threads

def mk():
    b = api.Batch(path, 1)
    it = seq["init"]
    b.set_initial_state(0, it["t"], it["quat"], it["pos"], it["vel"], it["bg"], it["ba"])
    return b

bs = [None if inside else mk() for _ in range(n_threads)]

def work(k):
    b = mk() if inside else bs[k]
    out[k] = b.replay(*pack)[0][0]

for rnd in range(2):
    ths = [threading.Thread(target=work, args=(k,)) for k in range(n_threads)]
    for t in ths: t.start()
    for t in ths: t.join()
    bad = {}
    for k in range(1, n_threads):
        d = np.abs(out[k] - out[0]).max(axis=1)
        if d.max() > 0:
            bad[k] = (int(np.argmax(d > 0)), float(d.max()))
    print("round", rnd, "handles made", "inside threads" if inside else "in the main thread", "-> differing:", bad)
    bs = [None if inside else mk() for _ in range(n_threads)]
