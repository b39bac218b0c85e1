"""GPU: the persistent frame handle (orcvio_frame_*: what bench.py times) gives exactly what the
oracle-checked one-shot entry (orcvio_snapshot_update, tests/test_gpu_stages.py) gives on the same window --
through the end-to-end call with host buffers (early prior factor + early direct-mode triangulation) and
through the resident load / run / fetch form, call after call."""
import numpy as np
import pytest

from orcvio_b200 import api, synth

pytestmark = pytest.mark.gpu

SIGMA2 = 1.6e-5
TRI = dict(cost_threshold=1e-3, init_final_dist_threshold=100.0)


def _frame(n_clones, flags=0):
    return api.Frame(n_clones, flags, SIGMA2, 0.95, -1.0, TRI["cost_threshold"], TRI["init_final_dist_threshold"])


def _same(a, b):
    np.testing.assert_array_equal(a["status"], b["status"])
    np.testing.assert_array_equal(a["gamma"], b["gamma"])
    np.testing.assert_array_equal(a["delta_x"], b["delta_x"])
    np.testing.assert_array_equal(a["P"], b["P"])
    np.testing.assert_array_equal(a["clones"], b["clones"])


@pytest.mark.parametrize("n_clones,n_feat,max_len,full", [
    (30, 2000, 6, False),
    (20, 300, 6, False),
    (12, 60, 6, True),            # long tracks: 128-thread Jacobian teams, wide A-form windows
    (20, 40, 6, False),           # fewer rows than columns
])
def test_frame_handle_equals_snapshot_entry(n_clones, n_feat, max_len, full):
    snap = synth.stress_snapshot(n_clones, n_feat, max_len, seed=21, full_tracks=full)
    ref = api.snapshot_update(snap, flags=0, noise_var=SIGMA2, translation_threshold=-1.0,
                              cost_threshold=TRI["cost_threshold"],
                              init_final_dist_threshold=TRI["init_final_dist_threshold"])
    assert ((ref["status"] & 2) != 0).sum() > 0
    fr = _frame(n_clones)
    inp = fr.prepare_inputs(snap)
    out = fr.update(inp)                     # end to end, host buffers
    _same(out, ref)
    out2 = fr.update(inp)                    # again on the same handle: nothing leaks from call to call
    _same(out2, ref)
    fr.load(snap)                            # resident form
    fr.run(2)
    _same(fr.fetch(), ref)
    other = synth.stress_snapshot(n_clones, max(n_feat // 2, 8), max_len, seed=22, full_tracks=full)
    ref_o = api.snapshot_update(other, flags=0, noise_var=SIGMA2, translation_threshold=-1.0,
                                cost_threshold=TRI["cost_threshold"],
                                init_final_dist_threshold=TRI["init_final_dist_threshold"])
    _same(fr.update(other), ref_o)           # a different frame through the same handle
    _same(fr.update(inp), ref)


def test_pose_cov_only_variant_matches_the_full_call():
    snap = synth.stress_snapshot(20, 300, 6, seed=31)
    fr = api.Frame(20, 0, 1.6e-5, 0.95, -1.0, 1e-3, 100.0)
    full = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in fr.update(snap).items()}
    lite = fr.update(snap, full_P=False)
    np.testing.assert_array_equal(lite["P_lead9"], full["P"][:9, :9])
    for key in ("delta_x", "status", "gamma", "clones"):
        np.testing.assert_array_equal(lite[key], full[key])
