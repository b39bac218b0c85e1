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
    print("  panel: start  upd  factor  solve  trail_done  sync   (cycles; deltas)")
    for k in list(range(0, min(npan, 4))) + list(range(npan // 2, npan // 2 + 2)) + [npan - 2, npan - 1]:
        r = p[k]
        print(f"  {k:3d}: start {r[0]-t0:7d} upd {r[1]-r[0]:5d} factor {r[2]-r[1]:5d} solve {r[3]-r[2]:5d} "
              f"trail_done {(r[4]-r[0]) if r[4] else 0:6d} total {r[5]-r[0]:6d}")
    print(f"  total cycles {p[npan-1,5]-t0}")
