"""Runs the same single-filter sequence twice and a batch twice: results must be bitwise equal
(no uninitialised reads, no order-dependent reductions)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from orcvio_b200 import api, montecarlo as mc
import helpers as H

ids = [0, 1, 2]
seqs = mc.make_sequences("unity", ids, 28, 60, dict(if_ZUPT_valid=0), n_landmarks=3000)


def single(s):
    vio = api.OrcVIO(H.write_cfg(s["cfg"])); assert vio.initialize()
    k = 0
    out = []
    for (t_img, feats) in s["frames"]:
        k1 = k
        while k1 < len(s["imu"]) and s["imu"][k1][0] <= t_img + 0.02: k1 += 1
        vio.push_imu(s["imu"][k:k1]); k = k1
        vio.processFeatures(t_img, feats)
        out.append((np.array(vio.state().p), vio.cov()))
    return out

a, b = single(seqs[1]), single(seqs[1])
print("single vs single bitwise:", all(np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) for x, y in zip(a, b)))
cfg = H.write_cfg(seqs[0]["cfg"])
r1, b1 = mc.run_local(cfg, seqs, ids)
r2, b2 = mc.run_local(cfg, seqs, ids)
print("batch vs batch bitwise:", np.array_equal(r1, r2), np.array_equal(b1.cov(1), b2.cov(1)))
# first frame at which batch filter 1 and the single filter differ
n = len(seqs)
bb = api.Batch(cfg, n)
for i, s in enumerate(seqs):
    it = s["init"]; bb.set_initial_state(i, it["t"], it["quat"], it["pos"], it["vel"], it["bg"], it["ba"])
cursor = [0] * n
for fi in range(28):
    t_img, feats, feat_off, imu, imu_off = mc.pack_frame(seqs, fi, cursor)
    used, pub = bb.process(t_img, feats, feat_off, imu, imu_off)
    for i in range(n): cursor[i] += int(used[i])
    dp = np.abs(np.array(bb.state(1).p) - a[fi][0]).max()
    dP = np.abs(bb.cov(1) - a[fi][1]).max()
    print(fi, "batch[1] vs single: dp %.2e dP %.2e" % (dp, dP), "N", bb.state(1).n_clones)
