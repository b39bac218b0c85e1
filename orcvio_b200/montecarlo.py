"""Multi-trajectory (Monte-Carlo / multi-sequence) replay: SURVEY 8e, BASELINE.json configs[4].

A filter is sequential frame to frame, so the path shards only across independent trajectories:
trajectory t goes to rank t mod world, every rank advances its trajectories in lock-step with
`orcvio_batch_*` (each kernel has a filter grid dimension), no collective touches the data path;
one all_gather of a fixed-size record per trajectory at the end (NCCL on the GPU box, gloo in the
CPU tests).  This module is host-side plumbing: sharding, input packing, record gathering.
"""
import numpy as np

from . import api, synth

RECORD = ("traj", "frames", "px", "py", "pz", "ate_m", "feature_updates", "ok")


def shard(n_traj, rank, world):
    """Trajectory indices owned by `rank` (round-robin: equal lengths -> balanced)."""
    return list(range(rank, n_traj, world))


def pack_frame(seqs, fi, imu_cursor):
    """Concatenate frame `fi` of every sequence into the CSR arrays orcvio_batch_process expects.
    imu_cursor[i] = first IMU sample of sequence i not yet consumed by the filter."""
    t_img = np.zeros(len(seqs))
    feats, feat_off, imus, imu_off = [], [0], [], [0]
    for i, s in enumerate(seqs):
        t, f = s["frames"][fi]
        t_img[i] = t
        feats.append(api.feats_array(f))
        feat_off.append(feat_off[-1] + len(f))
        imu = s["imu"]
        k0 = imu_cursor[i]
        k1 = k0
        while k1 < len(imu) and imu[k1][0] <= t + 0.02:
            k1 += 1
        imus.append(api.imu_array(imu[k0:k1]))
        imu_off.append(imu_off[-1] + (k1 - k0))
    return (t_img, np.concatenate(feats) if feats else np.zeros(0, dtype=api.FEAT_DTYPE),
            np.array(feat_off, dtype=np.int32), np.concatenate(imus), np.array(imu_off, dtype=np.int32))


def make_sequences(config, traj_ids, n_frames, feats_per_frame, overrides, n_landmarks=3000):
    return [synth.make_sequence(synth.SynthSpec(config=config, seed=int(t), n_frames=n_frames,
                                                feats_per_frame=feats_per_frame, overrides=overrides,
                                                n_landmarks=n_landmarks)) for t in traj_ids]


def run_local(cfg_path, seqs, traj_ids):
    """Advance the local trajectories in lock-step on the current device.  Returns the record
    matrix (len(seqs) x len(RECORD)) and the Batch (for timing / counters)."""
    n = len(seqs)
    b = api.Batch(cfg_path, n)
    for i, s in enumerate(seqs):
        it = s["init"]
        b.set_initial_state(i, it["t"], it["quat"], it["pos"], it["vel"], it["bg"], it["ba"])
    cursor = [0] * n
    n_frames = min(len(s["frames"]) for s in seqs)
    est = np.zeros((n, n_frames, 3))
    ok = np.ones(n)
    for fi in range(n_frames):
        t_img, feats, feat_off, imu, imu_off = pack_frame(seqs, fi, cursor)
        used, pub = b.process(t_img, feats, feat_off, imu, imu_off)
        for i in range(n):
            cursor[i] += int(used[i])
            ok[i] = min(ok[i], float(pub[i]))
            est[i, fi] = b.state(i).p[:]
    rec = np.zeros((n, len(RECORD)))
    for i, s in enumerate(seqs):
        gt = np.array([g[1] for g in s["gt"][:n_frames]])
        d = (est[i] - est[i, 0]) - (gt - gt[0])          # first-pose alignment (System.cpp:885-943)
        rec[i] = [traj_ids[i], n_frames, *est[i, -1], float(np.mean(np.linalg.norm(d, axis=1))), 0.0, ok[i]]
    rec[:, 6] = b.feature_updates() / max(n, 1)
    return rec, b


def gather_records(rec, n_traj, rank, world, device=None):
    """all_gather of the per-trajectory records; returns an (n_traj x len(RECORD)) matrix ordered
    by trajectory id on every rank.  Works on NCCL (device tensors) and gloo (CPU tensors)."""
    if world == 1:
        out = np.zeros((n_traj, len(RECORD)))
        out[rec[:, 0].astype(int)] = rec
        return out
    import torch
    import torch.distributed as dist
    per = (n_traj + world - 1) // world
    buf = torch.full((per, len(RECORD)), -1.0, dtype=torch.float64, device=device)
    if len(rec):
        buf[:len(rec)] = torch.as_tensor(rec, dtype=torch.float64, device=device)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = np.zeros((n_traj, len(RECORD)))
    for p in parts:
        a = p.cpu().numpy()
        a = a[a[:, 0] >= 0]
        out[a[:, 0].astype(int)] = a
    return out
