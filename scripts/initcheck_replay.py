import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orcvio_b200 import api, configs, synth, montecarlo as mc
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 70
ov = eval(sys.argv[2]) if len(sys.argv) > 2 else dict(max_features_in_one_grid=0)
seqs = [synth.make_sequence(synth.SynthSpec(config="euroc", seed=s, n_frames=n_frames, feats_per_frame=150, overrides=ov, n_landmarks=3000)) for s in (3,)]
path = os.path.join(tempfile.mkdtemp(), "cfg.yaml")
configs.write_yaml(path, seqs[0]["cfg"])
rec, info = mc.run_replay(path, seqs, [0], n_threads=1)
print("done", info["poses"][:, -1, :3])
