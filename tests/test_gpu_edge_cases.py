"""GPU: edge cases of the frame update through the C ABI -- empty and ragged inputs, nothing gated in, one
feature, observation lists too short for a projection, the largest window the library accepts, and rejected
(malformed) inputs.  Checked against the oracle where there is something to compare, and against the
"posterior == prior, state untouched" contract of removeLostFeatures (src/orcvio.cpp:2332-2336: return when no
feature is processed) otherwise."""
import numpy as np
import pytest

from orcvio_b200 import api, synth
import helpers as H

pytestmark = pytest.mark.gpu

SIGMA2 = 1.6e-5
TRI = dict(cost_threshold=1e-3, init_final_dist_threshold=100.0)


def _update(snap, sigma2=SIGMA2, tri=TRI, flags=0):
    return api.snapshot_update(snap, flags=flags, noise_var=sigma2, translation_threshold=-1.0,
                               cost_threshold=tri["cost_threshold"],
                               init_final_dist_threshold=tri["init_final_dist_threshold"])


def _subset(snap, keep):
    """Snapshot restricted to the features in `keep` (ragged CSR rebuilt)."""
    fo = np.asarray(snap["feat_off"])
    oc, oz, nfo = [], [], [0]
    for f in keep:
        oc.extend(snap["obs_clone"][fo[f]:fo[f + 1]])
        oz.extend(snap["obs_z"][fo[f]:fo[f + 1]])
        nfo.append(len(oc))
    s = dict(snap)
    s["feat_off"] = np.array(nfo, dtype=np.int32)
    s["obs_clone"] = np.array(oc, dtype=np.int32).reshape(-1)
    s["obs_z"] = np.array(oz, dtype=float).reshape(-1, 2)
    return s


def _prior_untouched(snap, out):
    np.testing.assert_array_equal(out["P"], np.asarray(snap["P"]))
    assert np.all(out["delta_x"] == 0.0)
    N = int(snap["n_clones"])
    np.testing.assert_array_equal(out["clones"][:, :9], np.asarray(snap["clone_R"]).reshape(N, 9))
    np.testing.assert_array_equal(out["clones"][:, 9:], np.asarray(snap["clone_p"]).reshape(N, 3))


def test_no_features_leaves_the_filter_alone():
    snap = _subset(synth.stress_snapshot(12, 8, 6, seed=1), [])
    out = _update(snap)
    assert out["status"].size == 0
    _prior_untouched(snap, out)


def test_nothing_gated_in_leaves_the_filter_alone():
    snap = synth.stress_snapshot(20, 120, 6, seed=2)
    # (a) every triangulation rejected by the cost threshold
    out = _update(snap, tri=dict(cost_threshold=1e-30, init_final_dist_threshold=100.0))
    assert np.all((out["status"] & 1) == 0) and np.all((out["status"] & 2) == 0)
    _prior_untouched(snap, out)
    # (b) every feature fails the chi-square gate: prior and measurement noise claimed 1e12 times tighter than
    # the residuals (gamma scales by 1e12)
    tight = dict(snap, P=np.asarray(snap["P"]) * 1e-12)
    out = _update(tight, sigma2=SIGMA2 * 1e-12)
    assert (out["status"] & 1).sum() > 100 and np.all((out["status"] & 2) == 0)
    _prior_untouched(tight, out)


def test_single_feature_matches_the_oracle():
    snap = _subset(synth.stress_snapshot(15, 40, 6, seed=3), [7])
    out = _update(snap)
    ref = H.oracle_snapshot_update(snap, 0, SIGMA2, tri=dict(translation_threshold=-1.0, **TRI))
    assert np.array_equal(out["status"], ref["status"]) and (out["status"] & 2).all()
    assert np.abs(out["delta_x"] - ref["delta_x"]).max() <= 1e-9 * np.abs(ref["delta_x"]).max()
    assert np.abs(out["P"] - ref["P"]).max() <= 1e-9 * np.abs(ref["P"]).max()


def test_ragged_tracks_including_too_short_ones():
    """A track with ONE observation has no left-nullspace row (2m - 3 < 1, nullspace_project_inplace_svd returns
    false, math_utils.hpp:290-297): it never reaches the update.  Two observations give one row (the pruning case,
    src/orcvio.cpp:2773) and take part like any longer track; everything that does must match the oracle."""
    base = synth.stress_snapshot(20, 60, 6, seed=4)
    fo = np.asarray(base["feat_off"])
    oc, oz, nfo = [], [], [0]
    for f in range(60):
        m = fo[f + 1] - fo[f]
        keep = m if f % 5 else min(m, 1 + (f // 5) % 2)        # every fifth track cut down to 1 or 2 observations
        oc.extend(base["obs_clone"][fo[f]:fo[f] + keep])
        oz.extend(base["obs_z"][fo[f]:fo[f] + keep])
        nfo.append(len(oc))
    snap = dict(base, feat_off=np.array(nfo, dtype=np.int32), obs_clone=np.array(oc, dtype=np.int32),
                obs_z=np.array(oz, dtype=float).reshape(-1, 2))
    out = _update(snap)
    m_all = np.diff(snap["feat_off"])
    short = m_all < 2
    assert short.sum() == 6 and (m_all == 2).sum() == 6 and np.all((out["status"][short] & 2) == 0)
    ref = H.oracle_snapshot_update(_subset(snap, np.flatnonzero(~short)), 0, SIGMA2,
                                   tri=dict(translation_threshold=-1.0, **TRI))
    assert np.array_equal(out["status"][~short] & 2, ref["status"] & 2)
    assert np.abs(out["delta_x"] - ref["delta_x"]).max() <= 1e-9 * np.abs(ref["delta_x"]).max()
    assert np.abs(out["P"] - ref["P"]).max() <= 1e-9 * np.abs(ref["P"]).max()


@pytest.mark.parametrize("full", [False, True])
def test_largest_window(full):
    """sw_size 31 is the largest window the library accepts (D = 208); full = every track spans the window
    (m = 31: 59 rows per feature, 186-column A-form windows, 128-thread Jacobian teams)."""
    snap = synth.stress_snapshot(31, 48 if full else 600, 6, seed=5, full_tracks=full)
    out = _update(snap)
    ref = H.oracle_snapshot_update(snap, 0, SIGMA2, tri=dict(translation_threshold=-1.0, **TRI))
    assert np.array_equal(out["status"] & 1, ref["status"] & 1)
    assert np.array_equal(out["status"] & 2, ref["status"] & 2) and (ref["status"] & 2).sum() > 0
    assert np.abs(out["delta_x"] - ref["delta_x"]).max() <= 1e-9 * np.abs(ref["delta_x"]).max()
    assert np.abs(out["P"] - ref["P"]).max() <= 1e-9 * np.abs(ref["P"]).max()
    assert np.abs(out["P"] - out["P"].T).max() == 0.0


def test_malformed_inputs_are_rejected():
    snap = synth.stress_snapshot(10, 20, 6, seed=6)
    bad = dict(snap, obs_clone=np.where(np.arange(len(snap["obs_clone"])) == 3, 10, snap["obs_clone"]).astype(np.int32))
    with pytest.raises(RuntimeError):
        _update(bad)                                           # clone index outside the window
    fr = api.Frame(10, 0, SIGMA2, 0.95, -1.0, TRI["cost_threshold"], TRI["init_final_dist_threshold"])
    with pytest.raises(RuntimeError):
        fr.update(bad)
    out = fr.update(snap)                                      # the handle survives a rejected call
    ref = _update(snap)
    np.testing.assert_array_equal(out["P"], ref["P"])
    with pytest.raises(RuntimeError):
        api.Frame(10).update(synth.stress_snapshot(12, 20, 6, seed=6))   # more clones than the handle's capacity
