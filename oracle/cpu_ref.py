"""CPU baseline driver (TEST / BENCH INFRASTRUCTURE, not product code).

time_frames() times the CPU restatement of the reference's frame update -- dense like the
reference -- on independent synthetic frames, one frame per worker.  It prefers the C++
restatement (oracle/cpu_ref.cpp -> oracle/_build/libcpu_ref.so, built by oracle/Makefile) and
falls back to the NumPy oracle (oracle/snapshot.py) when the shared object is absent.
Only bench.py's cpu_baseline / --impl reference legs and tests/ call this.
"""
import ctypes as C
import os
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcpu_ref.so")
_SO_O3 = os.path.join(_HERE, "_build", "libcpu_ref_o3.so")     # -O3 -march=x86-64-v3 build of the same file (timing only)


def _numpy_worker(args):
    n_clones, n_feat, max_len, sigma2, tri, seed = args
    from orcvio_b200 import synth
    from oracle.snapshot import oracle_snapshot_update
    snap = synth.stress_snapshot(n_clones, n_feat, max_len, seed=seed)
    t0 = time.perf_counter()
    out = oracle_snapshot_update(snap, 0, sigma2, tri=dict(translation_threshold=-1.0, **tri))
    return int(((out["status"] & 2) != 0).sum()), time.perf_counter() - t0


def load(variant=""):
    """The C++ restatement's shared object (variant "o3": the -O3 / AVX2 build), or None when it has not been built."""
    so = _SO_O3 if variant == "o3" else _SO
    if not os.path.exists(so):
        return None
    L = C.CDLL(so)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.cpu_ref_frame_update.restype = C.c_int
    L.cpu_ref_frame_update.argtypes = [dp, dp, C.c_int, dp, dp, dp, ip, ip, dp, C.c_int, C.c_int, C.c_double,
                                       C.c_double, C.c_double, C.c_double, C.c_double, dp, dp, ip, dp, dp, dp]
    L.cpu_ref_triangulate.restype = C.c_int
    L.cpu_ref_triangulate.argtypes = [dp, dp, ip, ip, dp, C.c_int, C.c_double, C.c_double, C.c_double, dp, ip, ip, dp]
    L.cpu_ref_chi2_quantile.restype = C.c_double
    L.cpu_ref_chi2_quantile.argtypes = [C.c_double, C.c_int]
    return L


def frame_update(snap, flags, sigma2, chi2_p=0.95, tri=None, variant=""):
    """Runs the C++ restatement on one frame; same outputs as oracle_snapshot_update."""
    L = load(variant)
    if L is None:
        raise RuntimeError("oracle/_build/libcpu_ref.so not built (make -C oracle)")
    tri = tri or {}
    N = int(snap["n_clones"])
    D = 22 + 6 * N
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    cR, cp = f64(snap["clone_R"]).reshape(N, 9), f64(snap["clone_p"]).reshape(N, 3)
    Rbc, tcb = f64(snap["R_b2c"]).reshape(9), f64(snap["t_c_b"]).reshape(3)
    P = np.asfortranarray(snap["P"], dtype=np.float64)
    fo, oc, oz = i32(snap["feat_off"]), i32(snap["obs_clone"]), f64(snap["obs_z"]).reshape(-1, 2)
    nf = len(fo) - 1
    out = dict(P=np.zeros((D, D), order="F"), delta_x=np.zeros(D), status=np.zeros(nf, dtype=np.int32),
               gamma=np.zeros(nf), positions=np.zeros((nf, 3)), clones=np.zeros((N, 12)))
    rc = L.cpu_ref_frame_update(dp(cR), dp(cp), N, dp(Rbc), dp(tcb), dp(P), ip(fo), ip(oc), dp(oz), nf, flags,
                                sigma2, chi2_p, tri.get("translation_threshold", -1.0),
                                tri.get("cost_threshold", 4.7673e-4), tri.get("init_final_dist_threshold", 5.0),
                                dp(out["P"]), dp(out["delta_x"]), ip(out["status"]), dp(out["gamma"]),
                                dp(out["positions"]), dp(out["clones"]))
    if rc != 0:
        raise RuntimeError(f"cpu_ref_frame_update failed: {rc}")
    return out


def _cpp_worker(args, variant=""):
    n_clones, n_feat, max_len, sigma2, tri, seed = args
    from orcvio_b200 import synth
    snap = synth.stress_snapshot(n_clones, n_feat, max_len, seed=seed)
    t0 = time.perf_counter()
    out = frame_update(snap, 0, sigma2, tri=dict(translation_threshold=-1.0, **tri), variant=variant)
    return int(((out["status"] & 2) != 0).sum()), time.perf_counter() - t0


def _cpp_worker_o3(args):
    return _cpp_worker(args, "o3")


def time_frames(n_clones, n_feat, max_len, sigma2, tri, n_threads, repeats=1, seed0=0, variant=""):
    """(features gated in, wall seconds of the slowest worker chain, kind)."""
    have_cpp = os.path.exists(_SO_O3 if variant == "o3" else _SO)
    worker = (_cpp_worker_o3 if variant == "o3" else _cpp_worker) if have_cpp else _numpy_worker
    jobs = [(n_clones, n_feat, max_len, sigma2, tri, seed0 + k) for k in range(n_threads * repeats)]
    feats, per_worker = 0, 0.0
    with ProcessPoolExecutor(max_workers=n_threads) as ex:
        t0 = time.perf_counter()
        res = list(ex.map(worker, jobs))
        wall = time.perf_counter() - t0
    feats = sum(r[0] for r in res)
    # the workers run concurrently: elapsed = the per-frame compute times laid on n_threads cores.
    # Frame generation happens inside the workers but outside their timed sections, so use the
    # sum of timed sections / n_threads (perfect packing, favourable to the CPU).
    per_worker = sum(r[1] for r in res) / n_threads
    return feats, min(per_worker, wall), "port"
