"""Per-phase clock profile of the one-CTA Cholesky (csrc/chol.cuh) on the GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orcvio_b200 import api

rng = np.random.default_rng(0)
for (m, nx) in [(202, 0), (180, 17), (64, 0)]:
    B = rng.standard_normal((m, m + 20))
    A = B @ B.T + 1e-3 * np.eye(m)
    X = rng.standard_normal((nx, m)) if nx else None
    L, Xs, prof, us = api.chol_probe(A, X, reps=20)
    Lr = np.linalg.cholesky(A)
    err = np.abs(np.tril(L) - Lr).max() / np.abs(Lr).max()
    msg = f"m={m} nx={nx}: {us:.1f} us, rel err L {err:.2e}"
    if nx:
        Xr = np.linalg.solve(Lr, X.T).T
        msg += f", X {np.abs(Xs - Xr).max() / np.abs(Xr).max():.2e}"
    print(msg)
    npan = (m + 7) // 8
    p = prof[:npan].astype(np.int64)
    t0 = p[0, 0]
    print("  step: chain [factor  wait_tile  solve+diag_upd  step]   bulk warp 0 [seen_block  solve  col_q+1  bulk_q+2]   (cycles)")
    for k in list(range(0, min(npan, 4))) + list(range(npan // 2, npan // 2 + 2)) + [npan - 2, npan - 1]:
        r = p[k]
        nxt = p[k + 1, 0] if k + 1 < npan else r[3]
        print(f"  {k:3d}: start {r[0]-t0:7d} | factor {r[1]-r[0]:5d} wait {max(r[2]-r[1],0):5d} solve_upd {r[3]-max(r[2],r[1]):5d} step {nxt-r[0]:5d}"
              f" | seen {r[4]-r[1]:5d} solve {r[5]-r[4]:5d} col {r[6]-r[5]:5d} bulk {r[7]-r[6]:5d}")
    print(f"  total cycles {max(p[npan-1,3], p[npan-1,7])-t0}")
