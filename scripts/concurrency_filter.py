"""The same sequence through N single-filter handles on N host threads: where do they first disagree?"""
import os, sys, threading, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orcvio_b200 import api, configs, synth
n_threads = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ov = eval(sys.argv[3]) if len(sys.argv) > 3 else dict(max_features_in_one_grid=0)
seq = synth.make_sequence(synth.SynthSpec(config="euroc", seed=3, n_frames=n_frames, feats_per_frame=150, overrides=ov, n_landmarks=3000))
path = os.path.join(tempfile.mkdtemp(), "cfg.yaml")
configs.write_yaml(path, seq["cfg"])
logs = [[] for _ in range(n_threads)]

def work(k):
    vio = api.OrcVIO(path)
    assert vio.initialize()
    c = 0
    for (t_img, f) in seq["frames"]:
        c1 = c
        while c1 < len(seq["imu"]) and seq["imu"][c1][0] <= t_img + 0.02:
            c1 += 1
        vio.push_imu(seq["imu"][c:c1])
        c = c1
        vio.processFeatures(t_img, f)
        st = vio.state()
        ids, ph, status, gamma = vio.candidate_log()
        P = vio.cov()
        logs[k].append(dict(p=np.array(st.p), R=np.array(st.R), ids=ids.copy(), ph=ph.copy(), status=status.copy(), gamma=gamma.copy(), P=P))

ths = [threading.Thread(target=work, args=(k,)) for k in range(n_threads)]
for t in ths: t.start()
for t in ths: t.join()
ref = logs[0]
for k in range(1, n_threads):
    for fi in range(n_frames):
        a, b = ref[fi], logs[k][fi]
        diffs = []
        for key in ("ids", "ph", "status", "gamma", "P", "p", "R"):
            if a[key].shape != b[key].shape or not np.array_equal(a[key], b[key]):
                if a[key].shape == b[key].shape:
                    d = np.abs(a[key].astype(float) - b[key].astype(float))
                    diffs.append((key, float(d.max()), int(np.argmax(d))))
                else:
                    diffs.append((key, "shape", a[key].shape, b[key].shape))
        if diffs:
            print(f"thread {k}: first disagreement at frame {fi}: {diffs}")
            if "gamma" in [d[0] for d in diffs]:
                d = np.abs(a["gamma"] - b["gamma"])
                j = np.flatnonzero(d > 0)
                print("    gamma differs for candidates", j[:10], "phase", a["ph"][j[:10]], "status", a["status"][j[:10]], b["status"][j[:10]])
            break
    else:
        print(f"thread {k}: identical")
