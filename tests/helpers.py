"""Shared test helpers: build oracle objects from synthetic snapshots / sequences."""
import os
import tempfile
from types import SimpleNamespace

import numpy as np

from oracle import feature as ofeat
from oracle import mathutils as mu
from oracle.filter import OracleVIO, Feature, Clone
from orcvio_b200 import configs, synth

FL_LARVIO, FL_LEFT, FL_DISCARD = 1, 2, 4


def write_cfg(cfg):
    path = os.path.join(tempfile.mkdtemp(prefix="orcvio_cfg_"), "cfg.yaml")
    configs.write_yaml(path, cfg)
    return path


from oracle.snapshot import oracle_from_snapshot, oracle_snapshot_update  # noqa: E402,F401


def run_oracle_sequence(seq, overrides=None):
    """Runs the oracle filter over a synthetic sequence; yields it after every frame."""
    path = write_cfg(seq["cfg"])
    vio = OracleVIO(path, overrides)
    assert vio.initialize()
    imu = [(r[0], r[1:4].copy(), r[4:7].copy()) for r in seq["imu"]]
    buf, k = [], 0
    for (t_img, feats) in seq["frames"]:
        while k < len(imu) and imu[k][0] <= t_img + 0.02:
            buf.append(imu[k])
            k += 1
        msg = (t_img, [tuple([int(f[0])] + list(f[1:])) for f in feats])
        vio.processFeatures(msg, buf)
        yield vio


def ate_first_pose_aligned(p_est, p_ref):
    """Mean position error after aligning the first pose (the reference's logger metric,
    ros_wrapper/src/orcvio/src/System.cpp:885-943, translation part)."""
    p_est = np.asarray(p_est)
    p_ref = np.asarray(p_ref)
    return float(np.mean(np.linalg.norm((p_est - p_est[0]) - (p_ref - p_ref[0]), axis=1)))
