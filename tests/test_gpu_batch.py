"""GPU: the multi-trajectory batch (orcvio_batch_*) advances B independent filters in lock-step
and gives, per filter, exactly what a single-filter handle gives on the same inputs."""
import numpy as np
import pytest

from orcvio_b200 import api, montecarlo as mc
import helpers as H

pytestmark = pytest.mark.gpu


def test_batch_equals_single_filters():
    ids = [0, 1, 2, 3, 4]
    ov = dict(if_ZUPT_valid=0)
    seqs = mc.make_sequences("unity", ids, 28, 60, ov, n_landmarks=3000)
    cfg = H.write_cfg(seqs[0]["cfg"])
    rec, b = mc.run_local(cfg, seqs, ids)
    assert np.all(rec[:, 7] == 1.0)
    assert b.kernel_launches() > 0 and b.feature_updates() > 100
    for i, s in enumerate(seqs):
        vio = api.OrcVIO(H.write_cfg(s["cfg"]))
        assert vio.initialize()
        k = 0
        for (t_img, feats) in s["frames"]:
            k1 = k
            while k1 < len(s["imu"]) and s["imu"][k1][0] <= t_img + 0.02:
                k1 += 1
            vio.push_imu(s["imu"][k:k1])
            k = k1
            assert vio.processFeatures(t_img, feats)
        st, sb = vio.state(), b.state(i)
        np.testing.assert_allclose(np.array(sb.p), np.array(st.p), rtol=0, atol=1e-12)
        np.testing.assert_allclose(np.array(sb.R), np.array(st.R), rtol=0, atol=1e-12)
        assert sb.n_clones == st.n_clones
        P1, P2 = vio.cov(), b.cov(i)
        assert np.abs(P1 - P2).max() <= 1e-12 * np.abs(P1).max()
        assert rec[i, 5] < 0.5                       # ATE vs synthetic ground truth [m]
    out = mc.gather_records(rec, len(ids), 0, 1)
    np.testing.assert_array_equal(out[:, 0], np.arange(len(ids)))
