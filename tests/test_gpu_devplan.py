"""The tile plan and the block pattern built ON THE DEVICE (csrc/devplan.cu) against the host statement of the same
layout (csrc/plan.cpp, csrc/rcm.cpp through mmba_plan_create / mmba_host_rcm_pattern): bit for bit."""
import numpy as np
import pytest

from meatmodeler_b200 import _capi, synth

pytestmark = pytest.mark.gpu


def _shuffled(prob, seed):
    rng = np.random.default_rng(seed)
    perm = rng.permutation(len(prob.uv))
    prob.uv, prob.cam_idx, prob.pt_idx = prob.uv[perm], prob.cam_idx[perm], prob.pt_idx[perm]
    return prob


def _with_duplicates_and_gaps(prob):
    """two observations of one point by the same camera, a 100-observation track and points nobody observes"""
    nc, npts, nobs = prob.sizes
    fi, pi, uv = prob.cam_idx.copy(), prob.pt_idx.copy(), prob.uv.copy()
    fi = np.concatenate((fi, fi[:7], np.arange(100) % nc))
    pi = np.concatenate((pi, pi[:7], np.full(100, 3)))
    uv = np.concatenate((uv, uv[:7] + 0.25, np.random.default_rng(1).normal(500, 50, (100, 2))))
    pi = np.where(pi >= npts - 5, pi - 7, pi)        # the last five points lose their observations
    prob.cam_idx, prob.pt_idx, prob.uv = fi, pi, uv
    return prob


CASES = {
    "windowed": lambda: synth.make_problem(40, 3000, 24000, seed=21, hard=True),
    "windowed_shuffled": lambda: _shuffled(synth.make_problem(40, 3000, 24000, seed=21, hard=True), 3),
    "random": lambda: _shuffled(synth.make_problem(60, 1500, 9000, seed=7, hard=True, windowed=False), 4),
    "ragged": lambda: synth.make_problem(300, 900, 4000, seed=9, windowed=False),
    "two_obs": lambda: synth.make_problem(12, 500, 1000, seed=10),
    "ring_wrap": lambda: synth.make_problem(200, 5000, 100000, seed=2, hard=True),
    "dups_gaps": lambda: _with_duplicates_and_gaps(synth.make_problem(150, 2000, 12000, seed=5, windowed=False)),
    "one_tile": lambda: synth.make_problem(5, 20, 60, seed=1),
    "c1": lambda: synth.make_config("C1", hard=True),
    "many_cams": lambda: synth.make_problem(1500, 20000, 90000, seed=8),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_device_plan_equals_host_plan(name):
    prob = CASES[name]()
    ext, K, pts, uv, fi, pi = prob.args()
    nc, npts = len(ext), len(pts)
    host = _capi.plan(nc, npts, fi, pi)
    with _capi.Engine() as eng:
        eng.set_problem(nc, npts, K, fi, pi, uv)
        dev = eng.plan()
        for k in ("n_tiles", "n_obs_local", "n_points_local", "point_begin", "point_end", "max_tile_cams", "n_slots"):
            assert dev[k] == host[k], k
        np.testing.assert_array_equal(dev["point_perm"], host["point_perm"])
        np.testing.assert_array_equal(dev["obs_perm"], host["obs_perm"])
        np.testing.assert_array_equal(dev["tile_cams"], host["tile_cams"])
        # raw tile records: header (pt0, npts, ncams, nobs, nruns, pair_mode, npairs, pad) + the five slot tables
        hdr_d = dev["meta"][:, :32].view(np.int32)
        hdr_h = host["meta"][:, :32].view(np.int32)
        np.testing.assert_array_equal(hdr_d, hdr_h)
        np.testing.assert_array_equal(dev["meta"], host["meta"])
        try:
            pat = eng.rcm_pattern()
        except _capi.MmbaError:
            pat = None
    rowptr, cols, nnz_full, total_pairs = _capi.host_rcm_pattern(nc, npts, fi, pi)
    if pat is not None:
        np.testing.assert_array_equal(pat["up_rowptr"], rowptr)
        np.testing.assert_array_equal(pat["up_cols"], cols)
        assert pat["nnz_full"] == nnz_full and pat["total_pairs"] == total_pairs


def test_device_plan_full_size_c2():
    prob = synth.make_config("C2", hard=True)
    ext, K, pts, uv, fi, pi = prob.args()
    nc, npts = len(ext), len(pts)
    host = _capi.plan(nc, npts, fi, pi)
    with _capi.Engine() as eng:
        eng.set_problem(nc, npts, K, fi, pi, uv)
        dev = eng.plan()
        pat = eng.rcm_pattern()
    assert dev["n_tiles"] == host["n_tiles"]
    np.testing.assert_array_equal(dev["point_perm"], host["point_perm"])
    np.testing.assert_array_equal(dev["obs_perm"], host["obs_perm"])
    np.testing.assert_array_equal(dev["meta"], host["meta"])
    np.testing.assert_array_equal(dev["tile_cams"], host["tile_cams"])
    rowptr, cols, nnz_full, total_pairs = _capi.host_rcm_pattern(nc, npts, fi, pi)
    np.testing.assert_array_equal(pat["up_rowptr"], rowptr)
    np.testing.assert_array_equal(pat["up_cols"], cols)
    assert pat["nnz_full"] == nnz_full and pat["total_pairs"] == total_pairs


def test_out_of_range_index_is_reported():
    prob = synth.make_problem(12, 500, 1000, seed=10)
    ext, K, pts, uv, fi, pi = prob.args()
    fi = fi.copy()
    fi[123] = len(ext)
    with _capi.Engine() as eng:
        with pytest.raises(_capi.MmbaError) as e:
            eng.set_problem(len(ext), len(pts), K, fi, pi, uv)
        assert e.value.code == -1 and "observation 123" in str(e.value)
        pi2 = pi.copy()
        pi2[77] = 2 ** 32 + 5          # must not alias point 5 after narrowing to 32 bits
        with pytest.raises(_capi.MmbaError) as e:
            eng.set_problem(len(ext), len(pts), K, prob.cam_idx, pi2, uv)
        assert "observation 77" in str(e.value)


def test_track_longer_than_a_tile_is_a_documented_error():
    nc = 300
    fi = np.concatenate((np.arange(257) % nc, [0, 1]))
    pi = np.concatenate((np.zeros(257, dtype=np.int64), [1, 1]))
    uv = np.zeros((259, 2))
    with _capi.Engine() as eng:
        with pytest.raises(_capi.MmbaError) as e:
            eng.set_problem(nc, 2, np.eye(3), fi, pi, uv)
        assert e.value.code == -5 and "point 0" in str(e.value)


def test_drop_in_reports_the_track_limit_as_value_error():
    """adjustPoints on a 257-observation track: a ValueError that names the limit (the reference has none; the
    drop-in fails loudly rather than solving a different problem)."""
    from meatmodeler_b200 import bundleAdjuster as mm
    nc = 300
    fi = np.concatenate((np.arange(257) % nc, [0, 1]))
    pi = np.concatenate((np.zeros(257, dtype=np.int64), [1, 1]))
    ext = np.tile(np.eye(4), (nc, 1, 1))
    with pytest.raises(ValueError, match="256 observations"):
        mm.adjustPoints(ext, np.eye(3), np.ones((2, 3)), np.zeros((259, 2)), fi, pi)
