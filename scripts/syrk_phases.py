"""Phase timeline of k_syrk on the frozen stress frame (ORCVIO_SYRK_DBG=1): per CTA, ns since the first CTA started."""
import ctypes as C, os, sys
os.environ["ORCVIO_SYRK_DBG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from orcvio_b200 import api
n_feat = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
snap = bench.make_frame(0, n_feat)
fr = api.Frame(bench.N_CLONES, 0, bench.NOISE_VAR, 0.95, -1.0, bench.TRI["cost_threshold"], bench.TRI["init_final_dist_threshold"])
fr.load(snap)
for _ in range(5):
    fr.run(1)
L = api.lib()
buf = np.zeros(1024 * 8, dtype=np.int64)
n = L.orcvio_syrk_debug(buf.ctypes.data_as(C.POINTER(C.c_longlong)), buf.size)
t = buf[:n].reshape(-1, 8).astype(np.float64)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
names = ["start", "prologue", "mainloop", "partial+atomic", "level1", "grp atomic", "level2", "emit"]
print("units", len(t))
for k in range(5):
    col = t[:, k]
    ok = col > 0
    if ok.any():
        print(f"{names[k]:15s} n={ok.sum():4d}  min {np.min(col[ok]-t0)/1e3:7.2f}  median {np.median(col[ok]-t0)/1e3:7.2f}  max {np.max(col[ok]-t0)/1e3:7.2f} us")
dur = (t[:, 2] - t[:, 1]) / 1e3
print("mainloop us per unit:", " ".join(f"{x:.1f}" for x in dur))
print("rows per unit:", " ".join(str(int(x)) for x in t[:, 5]))
print("pair (I*16+J) per unit:", " ".join(str(int(x)) for x in t[:, 6]))
print("jrow0[J] per unit:", " ".join(str(int(x)) for x in t[:, 7]))
