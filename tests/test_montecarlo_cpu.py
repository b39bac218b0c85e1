"""CPU tests of the multi-trajectory plumbing (world_size 2 over gloo): sharding covers every
trajectory once, the CSR packing is consistent, and the record all_gather reassembles the
per-trajectory records in trajectory order on every rank.  No GPU compute is involved."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from orcvio_b200 import montecarlo as mc


def test_shard_is_a_partition():
    for n, w in ((1024, 8), (10, 4), (3, 8), (7, 1)):
        owned = [mc.shard(n, r, w) for r in range(w)]
        flat = sorted(t for o in owned for t in o)
        assert flat == list(range(n))
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


def test_pack_frame_csr():
    seqs = mc.make_sequences("unity", [0, 1, 2], 3, 12, dict(if_ZUPT_valid=0), n_landmarks=800)
    cursor = [0, 0, 0]
    t_img, feats, feat_off, imu, imu_off = mc.pack_frame(seqs, 1, cursor)
    assert len(t_img) == 3 and feat_off[0] == 0 and imu_off[0] == 0
    assert feat_off[-1] == len(feats) and imu_off[-1] == len(imu)
    for i, s in enumerate(seqs):
        f = s["frames"][1][1]
        assert feat_off[i + 1] - feat_off[i] == len(f)
        np.testing.assert_array_equal(feats["id"][feat_off[i]:feat_off[i + 1]], f[:, 0].astype(np.uint64))
        seg = imu["t"][imu_off[i]:imu_off[i + 1]]
        assert np.all(np.diff(seg) > 0) and seg[-1] <= t_img[i] + 0.02


def _worker(rank, world, port, n_traj, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = mc.shard(n_traj, rank, world)
    rec = np.zeros((len(mine), len(mc.RECORD)))
    for k, t in enumerate(mine):
        rec[k] = [t, 40, 0.1 * t, -0.2 * t, 1.0, 1e-3 * (t + 1), 100 + t, 1.0]
    out = mc.gather_records(rec, n_traj, rank, world)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_records_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_traj, world = 7, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_traj, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        out = results[r]
        assert out.shape == (n_traj, len(mc.RECORD))
        np.testing.assert_array_equal(out[:, 0], np.arange(n_traj))
        np.testing.assert_allclose(out[:, 5], 1e-3 * (np.arange(n_traj) + 1))
    np.testing.assert_array_equal(results[0], results[1])


def test_pack_replay_layout():
    seqs = mc.make_sequences("unity", [0, 1, 2], 4, 12, dict(if_ZUPT_valid=0), n_landmarks=800, workers=2)
    t_img, feats, feat_off, imus = mc.pack_replay(seqs, 4)
    assert t_img.shape == (3, 4) and feat_off.shape == (3, 5) and len(feats) == len(imus) == 3
    for i, s in enumerate(seqs):
        assert feat_off[i, 0] == 0 and feat_off[i, -1] == len(feats[i])
        for f, (t, fr) in enumerate(s["frames"][:4]):
            assert t_img[i, f] == t
            np.testing.assert_array_equal(feats[i]["id"][feat_off[i, f]:feat_off[i, f + 1]], fr[:, 0].astype(np.uint64))
        assert len(imus[i]) == len(s["imu"])
    g = mc.gt_poses(seqs, 4)
    assert g.shape == (3, 4, 7) and np.allclose(np.linalg.norm(g[..., 3:], axis=-1), 1.0)
