#!/usr/bin/env python
"""bench.py -- feature updates / s of the OrcVIO filter-update hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], "stress frame"): one frozen frame with a 30-clone window
(D = 202) and 4096 features, reference-faithful track lengths m in [3, 6] (`max_track_len: 6`
in every shipped yaml; SURVEY 8d case 4a).  One step = one pass of the hot path over the frame:
per-feature LM triangulation, measurement Jacobians + left-nullspace projection + chi-square
gate, QR compression of the stacked H, FP64 EKF gain + covariance update (stages 1,2,4,5 of
BASELINE.json's north_star).  metric = features that pass the gate and update the filter per
second; `us_per_frame` is the same time per frame.

  value : kernel chain on the HBM-resident frame, CUDA events on the launching stream, L2
          flushed between timed iterations.
  e2e   : orcvio_frame_update() through the C ABI with HOST buffers every step (host work-list
          build + H2D + kernels + D2H of P / delta_x / gate decisions inside the timed region).
  N > 1 : the path shards only across independent trajectories: every rank owns its own frame
          (seed = rank), no data-path collective; NCCL gathers the per-rank counters.  Weak
          scaling; time = max over ranks.
  --impl reference : the CPU restatement of the reference algorithm (oracle/, dense like the
          reference) on the box's host cores, one independent frame per thread.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "feature_updates_per_sec"
UNIT = "features/s"
N_CLONES, N_FEATURES, MAX_TRACK = 30, 4096, 6
TARGET_FEATURES = 2000      # BASELINE.json north_star: "30-clone, 2000-feature frame runs under 200 us"
NOISE_VAR = 1.6e-5          # (2 x 0.002)^2: synthetic pixel noise of the KITTI-shaped generator
TRI = dict(cost_threshold=1e-3, init_final_dist_threshold=100.0)
# dram__bytes_read.sum + dram__bytes_write.sum of one k_syrk launch from the committed `ncu --set full`
# capture (profiles/); None until a capture of the current kernel exists
SYRK_NCU_TRAFFIC = 28_166_400          # 28.130 MB read + 36 KB written (profiles/r1_ncu_full_summary.csv): less than the
                                       # dense 8 M (n+1) B because the zero part of A's staircase is never read
WORKLOAD = f"stress frame: {N_CLONES}-clone window, {N_FEATURES} features, max_track_len {MAX_TRACK} (SURVEY 8d C4a)"


def make_frame(seed, n_feat=N_FEATURES):
    from orcvio_b200 import synth
    return synth.stress_snapshot(N_CLONES, n_feat, MAX_TRACK, seed=seed)


def algorithmic_bytes(snap, status, stage):
    """Compulsory HBM bytes of one launch of the dominant kernel (SURVEY 8d per-feature figures)."""
    fo = np.asarray(snap["feat_off"])
    m = np.diff(fo).astype(np.int64)
    N = int(snap["n_clones"])
    D = 22 + 6 * N
    r = 2 * m - 3
    c = 6 * m
    passed = (np.asarray(status) & 2) != 0
    per_frame = 8 * 24 * N + 8 * D * D            # clone records + covariance, read once
    if stage == "tri":
        return int((16 * m + 4 * m + 64 + 32 + 4).sum() + 8 * 24 * N)
    if stage == "jac_gate":
        # in: obs (16m) + clone idx (4m) + work record (64) + position (32);
        # out: gamma + status (12) and, for gated-in features, the compact r x (c+1) block
        return int((16 * m + 4 * m + 64 + 32 + 12).sum() + (8 * r * (c + 1))[passed].sum() + per_frame)
    if stage in ("qr_tiles", "qr_chain"):
        return int((8 * r * (c + 1))[passed].sum() + 8 * (6 * N) * (6 * N + 1))
    if stage == "update":
        return int(24 * D * D)
    return 0


def syrk_flops(snap, status):
    """Algorithmic flops of W = s^2 I + A^T A (k_syrk): one triangle of the (n+1) x (n+1) Gram
    matrix of the M gated rows of A = [H' L | r'], 2 flops per multiply-add  ->  M (n+1)(n+2).
    (The reference spends 2 M De^2 - 2/3 De^3 on the Householder QR of the same rows)"""
    m = np.diff(np.asarray(snap["feat_off"])).astype(np.int64)
    passed = (np.asarray(status) & 2) != 0
    M = int((2 * m - 3)[passed].sum())
    n1 = 6 * int(snap["n_clones"]) + 1
    return M * n1 * (n1 + 1), M


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for k, nm in enumerate(names):
                    if r[2 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=smax, reasons=sorted(reasons),
                    samples=len(sm))


def run_cpu_reference(n_threads, n_feat, repeats, seed0=0):
    """Times the CPU restatement of the reference algorithm on `n_threads` independent frames
    (one per thread).  Returns (features gated in, seconds, kind)."""
    from oracle import cpu_ref
    return cpu_ref.time_frames(N_CLONES, n_feat, MAX_TRACK, NOISE_VAR, TRI, n_threads, repeats, seed0)


def reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_feat = args.ref_features
    # warm-up + timed steps, each step = one bounded sample (one frame per thread)
    for _ in range(max(args.warmup, 0)):
        run_cpu_reference(cores, min(n_feat, 512), 1)
    feats, secs = 0, 0.0
    for k in range(args.steps):
        f, s, kind = run_cpu_reference(cores, n_feat, 1, seed0=100 * k)
        feats += f
        secs += s
    value = feats / secs
    sample = (f"{cores} independent frames per step (one per thread), each {N_CLONES} clones x {n_feat} features, "
              f"m in [3,{MAX_TRACK}]; dense per-feature H P H^T gate, Householder QR compression, dense EKF update")
    line = dict(metric=METRIC, value=value, unit=UNIT, impl="reference", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * secs / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=WORKLOAD, sample=sample),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind=kind, sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-features", type=int, default=N_FEATURES, help="features per frame of the CPU sample")
    ap.add_argument("--ref-repeats", type=int, default=6, help="frames per host thread in the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="skip the L2 flush (profiling runs only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from orcvio_b200 import api

    if not torch.cuda.is_available() or api.lib().orcvio_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: orcvio_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    api.lib().orcvio_set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    snap = make_frame(seed=rank)
    fr = api.Frame(N_CLONES, 0, NOISE_VAR, 0.95, -1.0, TRI["cost_threshold"], TRI["init_final_dist_threshold"])
    inp = fr.prepare_inputs(snap)
    out = fr.update(inp)                               # also the correctness anchor of this run
    n_pass = int(((out["status"] & 2) != 0).sum())
    n_valid = int(((out["status"] & 1) != 0).sum())
    fr.load(snap)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def l2_flush():
        if not args.no_flush:
            flush.zero_()
            torch.cuda.synchronize()

    # ---- value: HBM-resident frame, CUDA events around the kernel chain
    for _ in range(args.warmup):
        l2_flush()
        fr.run(1)
    _, stages = fr.run(5, stages=True)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = fr.kernel_launches()
    dev_us = 0.0
    for _ in range(args.steps):
        l2_flush()
        dev_us += fr.run(1)
    launches = fr.kernel_launches() - l0
    barrier()
    t_val = torch.tensor([dev_us * 1e-6], dtype=torch.float64, device=dev)

    # ---- e2e: host buffers through the C ABI, wall clock around synchronous calls
    for _ in range(args.warmup):
        fr.update(inp, out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fr.update(inp, out)
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- the north star's target frame (30 clones, 2000 features, target < 200 us), rank 0 only: same chain, same
    # timing method (CUDA events on the launching stream, L2 flushed), reported beside the headline workload
    target = None
    if rank == 0:
        snap_t = make_frame(seed=0, n_feat=TARGET_FEATURES)
        fr_t = api.Frame(N_CLONES, 0, NOISE_VAR, 0.95, -1.0, TRI["cost_threshold"], TRI["init_final_dist_threshold"])
        out_t = fr_t.update(fr_t.prepare_inputs(snap_t))
        fr_t.load(snap_t)
        for _ in range(args.warmup):
            l2_flush()
            fr_t.run(1)
        us_t, n_t = 0.0, max(20, min(args.steps, 100))
        for _ in range(n_t):
            l2_flush()
            us_t += fr_t.run(1)
        _, st_t = fr_t.run(5, stages=True)
        target = dict(workload=f"{N_CLONES}-clone window, {TARGET_FEATURES} features, max_track_len {MAX_TRACK}",
                      us_per_frame=us_t / n_t, target_us=200.0, gated_in=int(((out_t["status"] & 2) != 0).sum()),
                      stage_us=dict(tri=round(st_t["tri"], 2), jac_gate=round(st_t["jac_gate"], 2),
                                    aform_incl_prior_wait=round(st_t["qr_tiles"], 2),
                                    syrk_plus_chol_w_solve=round(st_t["qr_chain"], 2),
                                    pinfo_increment=round(st_t["update"], 2), total=round(st_t["total"], 2)))
        del fr_t

    counts = torch.tensor([float(n_pass), float(N_FEATURES)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_val, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    total_pass = counts[0].item()
    secs, secs_e2e = t_val.item(), t_e2e.item()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if peaks else "fallback 6650 GB/s"
        # stage split of the whitened-form path (api.Frame.STAGES keeps the generic names):
        #   qr_tiles = k_aform (+ wait for k_chol_prior), qr_chain = k_syrk + k_chol_w_solve, update = k_pinfo
        kt = fr.kernel_times()
        stage_named = dict(tri=stages["tri"], jac_gate=stages["jac_gate"], aform_incl_prior_wait=stages["qr_tiles"],
                           syrk=kt["syrk"], chol_w_solve=stages["qr_chain"] - kt["syrk"], pinfo_increment=stages["update"],
                           chol_prior_second_stream=kt["chol_prior"], total=stages["total"])
        # roofline kernel: k_syrk, the one kernel of the chain with GEMM-sized work (the FP64 tensor-core
        # compression GEMM of the north star); the other kernels are latency chains
        dfma_peak, dmma_peak = api.fp64_peak()
        flops, M_rows = syrk_flops(snap, out["status"])
        achieved = flops / (kt["syrk"] * 1e-6) / 1e12
        jac_bytes = algorithmic_bytes(snap, out["status"], "jac_gate")
        h2d = sum(inp[k].nbytes for k in ("clone_R", "clone_p", "Rbc", "tcb", "P", "feat_off", "obs_clone", "obs_z"))
        d2h = sum(out[k].nbytes for k in ("P", "delta_x", "status", "gamma", "clones"))
        line = dict(
            metric=METRIC, value=total_pass * args.steps / secs, unit=UNIT, n_gpus=world, steps=args.steps,
            warmup=args.warmup, ms_per_step=1e3 * secs / args.steps, higher_is_better=True, scaling="weak",
            vs_baseline=None, dtype="f64", data="synthetic",
            config=dict(workload=WORKLOAD, n_clones=N_CLONES, state_dim=22 + 6 * N_CLONES,
                        features_per_frame=N_FEATURES, triangulated_ok=n_valid, gated_in=n_pass,
                        frames_per_step_per_gpu=1, sharding="independent frames (trajectories) per rank, no collective",
                        l2="flushed between timed iterations (256 MiB memset)" if not args.no_flush else "not flushed",
                        timing="CUDA events on the launching stream per iteration, max over ranks"),
            us_per_frame=1e6 * secs / args.steps,
            stage_us={k: round(v, 2) for k, v in stage_named.items()},
            e2e=dict(value=total_pass * args.steps / secs_e2e, unit=UNIT, h2d_bytes_per_step=int(h2d),
                     d2h_bytes_per_step=int(d2h), us_per_frame=1e6 * secs_e2e / args.steps),
            gpu_launches=int(launches),
            roofline=dict(bound="tensor", kernel="k_syrk (W = s^2 I + A^T A, staircase-sparse split-K, mma.sync m8n8k4 f64)", achieved=achieved,
                          peak=dmma_peak, unit="TFLOP/s", frac=achieved / dmma_peak, traffic=SYRK_NCU_TRAFFIC,
                          algorithmic_flops=flops, gated_rows=M_rows, kernel_us=kt["syrk"],
                          peak_source="FP64 DMMA peak measured live on this GPU (orcvio_fp64_peak: mma.sync m8n8k4 "
                                      "micro-kernel; MEASURED_PEAKS.json carries only bf16 / HBM)",
                          fp64_dfma_peak=dfma_peak),
            roofline_hbm=dict(bound="hbm", kernel="k_jac_gate", achieved=jac_bytes / (stages["jac_gate"] * 1e-6) / 1e9,
                              peak=peak, unit="GB/s", frac=jac_bytes / (stages["jac_gate"] * 1e-6) / 1e9 / peak,
                              algorithmic_bytes=jac_bytes, kernel_us=stages["jac_gate"], peak_source=peak_src,
                              note="one frame's working set is L2-resident: latency-bound, not HBM-bound (SURVEY 8d)"),
            north_star_frame=target,
            clocks=clocks)
        if not args.no_cpu_baseline:
            try:
                cores = os.cpu_count() or 1
                f, s, kind = run_cpu_reference(cores, args.ref_features, args.ref_repeats)
                line["cpu_baseline"] = dict(
                    value=f / s, unit=UNIT, cores=cores, kind=kind,
                    sample=f"{cores * args.ref_repeats} independent frames ({args.ref_repeats} per host thread), each "
                           f"{N_CLONES} clones x {args.ref_features} features, m in [3,{MAX_TRACK}]; "
                           f"{s * cores:.1f} core-seconds of CPU work")
            except Exception as e:       # the baseline is a reported number, never the product path
                line["cpu_baseline"] = dict(value=None, unit=UNIT, cores=0, kind="port", sample=f"failed: {e}")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
