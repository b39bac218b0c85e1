#!/usr/bin/env python
"""BASELINE.json configs[4]: Monte-Carlo replay of independent synthetic EuRoC-shaped trajectories, sharded
over the GPUs of one box (trajectory t -> rank t mod world), every rank advancing its trajectories in lock-step
through orcvio_batch_process; NCCL only gathers the per-trajectory records at the end.

    python scripts/bench_montecarlo.py --traj 256 --frames 60
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        scripts/bench_montecarlo.py --traj 256 --frames 60

One JSON line on rank 0: trajectory-frames/s and feature updates/s of the whole job (max time over ranks),
ATE statistics against the synthetic ground truth.  Not the headline bench (bench.py): a side measurement.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--traj", type=int, default=256)
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--feats", type=int, default=150, help="visible features per frame")
    ap.add_argument("--config", default="euroc")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    from orcvio_b200 import api, configs, montecarlo as mc
    torch.cuda.set_device(local)
    api.lib().orcvio_set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mine = mc.shard(args.traj, rank, world)
    t0 = time.perf_counter()
    seqs = mc.make_sequences(args.config, mine, args.frames, args.feats, dict(if_ZUPT_valid=0, max_features_in_one_grid=0), n_landmarks=3000)
    t_gen = time.perf_counter() - t0
    import tempfile
    path = os.path.join(tempfile.mkdtemp(prefix="orcvio_mc_"), "cfg.yaml")
    configs.write_yaml(path, seqs[0]["cfg"])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # the lock-step loop of montecarlo.run_local with the library call timed apart from the Python packing
    n = len(seqs)
    b = api.Batch(path, n)
    for i, sq in enumerate(seqs):
        it = sq["init"]
        b.set_initial_state(i, it["t"], it["quat"], it["pos"], it["vel"], it["bg"], it["ba"])
    cursor = [0] * n
    est = np.zeros((n, args.frames, 3))
    ok = np.ones(n)
    t_lib = 0.0
    t0 = time.perf_counter()
    for fi in range(args.frames):
        packed = mc.pack_frame(seqs, fi, cursor)
        t1 = time.perf_counter()
        used, pub = b.process(*packed)
        t_lib += time.perf_counter() - t1
        for i in range(n):
            cursor[i] += int(used[i])
            ok[i] = min(ok[i], float(pub[i]))
            est[i, fi] = b.state(i).p[:]
    torch.cuda.synchronize()
    rec = np.zeros((n, len(mc.RECORD)))
    for i, sq in enumerate(seqs):
        gt = np.array([g[1] for g in sq["gt"][:args.frames]])
        d = (est[i] - est[i, 0]) - (gt - gt[0])
        rec[i] = [mine[i], args.frames, *est[i, -1], float(np.mean(np.linalg.norm(d, axis=1))), 0.0, ok[i]]
    rec[:, 6] = b.feature_updates() / max(n, 1)
    secs = torch.tensor([time.perf_counter() - t0, t_lib], dtype=torch.float64, device=dev)
    counts = torch.tensor([float(b.feature_updates()), float(b.kernel_launches())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(secs, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    out = mc.gather_records(rec, args.traj, rank, world, device=dev)
    if rank == 0:
        s, s_lib = secs[0].item(), secs[1].item()
        print(json.dumps(dict(
            workload=f"Monte-Carlo replay: {args.traj} independent {args.config}-shaped trajectories x {args.frames} "
                     f"frames, ~{args.feats} features per frame, sharded round-robin over {world} GPU(s)",
            n_gpus=world, trajectories=args.traj, frames=args.frames, seconds_in_library=s_lib,
            seconds_incl_python_packing=s,
            trajectory_frames_per_sec=args.traj * args.frames / s_lib, feature_updates_per_sec=counts[0].item() / s_lib,
            feature_updates=counts[0].item(), kernel_launches=counts[1].item(),
            all_published=bool(np.all(out[:, 7] == 1.0)), ate_m_mean=float(out[:, 5].mean()),
            ate_m_max=float(out[:, 5].max()), sequence_generation_s_per_rank=t_gen,
            timing="wall clock inside orcvio_batch_process (C++ bookkeeping of every filter + uploads + kernels + "
                   "mirror read-back per frame), max over ranks; the Python packing of the synthetic inputs is "
                   "reported apart")), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
