"""CPU-side checks of libmmba.so: it loads, exports every symbol of include/mmba.h, its host-only
functions agree with the scipy routines they replace, and the tile plan is a valid reordering.
No compute entry point is called here (no GPU in this suite)."""
import ctypes
import os
import re

import numpy as np
import pytest
from scipy.optimize._lsq import common as sc

from meatmodeler_b200 import _capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mmba.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mmba_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 25
    lib = ctypes.CDLL(_capi._LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"libmmba.so lacks {n}"
    assert sorted(_capi.SIGNATURES) == names      # the ctypes layer binds exactly the header
    assert _capi.lib().mmba_version() == 100


def test_struct_layout_matches_header():
    # sizes of the C structs as laid out by the compiler (x86-64 SysV)
    assert ctypes.sizeof(_capi.Options) == 16 + 24 + 8 + 8 + 8 + 128 + 8 + 8 + 8
    assert ctypes.sizeof(_capi.Result) == 24 + 24 + 8 + 8 + 8
    assert ctypes.sizeof(_capi.IterLog) == 9 * 8
    opt = _capi.default_options()
    assert (opt.ftol, opt.xtol, opt.gtol, opt.nranks, opt.max_nfev) == (1e-4, 1e-8, 1e-8, 1, 0)
    assert opt.schur_mode == _capi.SCHUR_AUTO and opt.pcg_atol == 1e-7 and opt.pcg_ktol == 1.23e-6


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_capi.MmbaError) as e:
        _capi.Engine()
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


# ---- scalar trust-region helpers vs scipy (common.py) -------------------------------------------

def test_min_quadratic_1d_matches_scipy():
    rng = np.random.default_rng(0)
    for _ in range(200):
        a, b = rng.normal(size=2) * 10 ** rng.uniform(-3, 3)
        lb, ub = np.sort(rng.normal(size=2))
        t, y = _capi.host_min_quadratic_1d(a, b, lb, ub)
        ts, ys = sc.minimize_quadratic_1d(a, b, lb, ub)
        assert y == pytest.approx(ys, rel=1e-14, abs=1e-300)
        assert t == pytest.approx(ts, rel=1e-14, abs=1e-300)
    # scipy's own known answers (scipy/optimize/tests/test_lsq_common.py:150-194)
    assert _capi.host_min_quadratic_1d(5, 2, -1, 1)[0] == pytest.approx(-0.2)
    assert _capi.host_min_quadratic_1d(5, 2, -1, 1)[1] == pytest.approx(5 * 0.04 - 0.4)
    assert _capi.host_min_quadratic_1d(-5, 2, -1, 1) == pytest.approx((-1, -7))


def test_update_tr_radius_matches_scipy():
    rng = np.random.default_rng(1)
    for _ in range(300):
        delta, sn = np.abs(rng.normal(size=2)) + 1e-3
        actual, pred = rng.normal(size=2)
        if rng.random() < 0.1:
            pred = 0.0
            if rng.random() < 0.5:
                actual = 0.0
        hit = bool(rng.random() < 0.5)
        assert _capi.host_update_tr_radius(delta, actual, pred, sn, hit) == pytest.approx(
            sc.update_tr_radius(delta, actual, pred, sn, hit), rel=1e-15)


def test_check_termination_matches_scipy():
    rng = np.random.default_rng(2)
    for _ in range(300):
        dF, F, dx, xn = 10 ** rng.uniform(-12, 2, size=4)
        ratio = rng.uniform(-0.5, 1.5)
        assert _capi.host_check_termination(dF, F, dx, xn, ratio, 1e-4, 1e-8) == sc.check_termination(
            dF, F, dx, xn, ratio, 1e-4, 1e-8)


def _model(B, g, p):
    return 0.5 * p @ B @ p + g @ p


def test_tr2d_matches_scipy():
    """solve_trust_region_2d has no unit test even in scipy: known answers are produced
    by calling scipy itself.  The minimiser may be non-unique in degenerate cases, so compare the
    model value, feasibility and the Newton flag; the point itself where it is well conditioned."""
    rng = np.random.default_rng(3)
    n_cmp = 0
    for k in range(600):
        A = rng.normal(size=(3, 2)) * 10 ** rng.uniform(-2, 2)
        B = A.T @ A if k % 3 else A.T @ A - np.eye(2) * rng.uniform(0, 2)   # some indefinite
        B = 0.5 * (B + B.T)
        g = rng.normal(size=2) * 10 ** rng.uniform(-2, 2)
        delta = 10 ** rng.uniform(-3, 2)
        p, newton = _capi.host_tr2d(B, g, delta)
        ps, ns = sc.solve_trust_region_2d(B, g, delta)
        assert newton == ns
        assert np.linalg.norm(p) <= delta * (1 + 1e-12)
        scale = max(abs(_model(B, g, ps)), 1e-300)
        assert _model(B, g, p) <= _model(B, g, ps) + 1e-9 * scale
        if not ns:
            assert np.linalg.norm(p) == pytest.approx(delta, rel=1e-12)
        ev = np.linalg.eigvalsh(B)
        if ev[1] - ev[0] > 1e-3 * abs(ev[1]):
            assert np.allclose(p, ps, rtol=1e-6, atol=1e-9 * delta)
            n_cmp += 1
    assert n_cmp > 300


# ---- tile plan ----------------------------------------------------------------------------------

def _check_plan(nc, npts, fi, pi, nranks):
    seen_obs = np.zeros(len(fi), dtype=int)
    seen_pts = np.zeros(npts, dtype=int)
    perm0 = None
    for rank in range(nranks):
        pl = _capi.plan(nc, npts, fi, pi, rank, nranks)
        perm = pl["point_perm"]
        if perm0 is None:
            perm0 = perm
            assert sorted(perm) == list(range(npts))
        np.testing.assert_array_equal(perm, perm0)          # identical on every rank
        live = pl["obs_perm"] >= 0
        obs = pl["obs_perm"][live]
        assert len(obs) == pl["n_obs_local"]
        seen_obs[obs] += 1
        # every slot resolves to the camera / point of its observation through the tile tables
        np.testing.assert_array_equal(pl["slot_cam"][live], fi[obs])
        np.testing.assert_array_equal(perm[pl["point_begin"] + pl["slot_point"][live]], pi[obs])
        seen_pts[perm[pl["point_begin"]:pl["point_end"]]] += 1
        # a point never straddles two tiles
        tile = np.nonzero(live)[0] // pl["tile_obs"]
        local_pt = pl["slot_point"][live]
        first = {}
        for t, q in zip(tile, local_pt):
            assert first.setdefault(q, t) == t
        assert pl["max_tile_cams"] <= pl["tile_obs"]
    assert np.all(seen_obs == 1) and np.all(seen_pts == 1)


@pytest.mark.parametrize("nranks", [1, 2, 3])
def test_plan_is_a_partition(nranks):
    prob = synth.make_problem(30, 700, 5000, seed=3, windowed=False)
    fi, pi = prob.cam_idx, prob.pt_idx
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(fi))                      # any observation order is accepted
    _check_plan(30, 700, fi[perm], pi[perm], nranks)


def test_plan_balances_observations():
    prob = synth.make_problem(50, 4000, 30000, seed=4)
    loads = [_capi.plan(50, 4000, prob.cam_idx, prob.pt_idx, r, 4)["n_obs_local"] for r in range(4)]
    assert sum(loads) == 30000
    assert max(loads) - min(loads) <= 2 * 8 + 1          # within a track length or two


def test_plan_edge_cases():
    # unobserved points, a single observation per point, more ranks than points
    fi = np.array([0, 1, 1, 0], dtype=np.int64)
    pi = np.array([3, 3, 0, 5], dtype=np.int64)
    _check_plan(2, 7, fi, pi, 1)
    _check_plan(2, 7, fi, pi, 5)
    with pytest.raises(_capi.MmbaError) as e:
        _capi.plan(2, 7, np.array([0, 2], dtype=np.int64), np.array([0, 1], dtype=np.int64))
    assert e.value.code == -1
    # a track longer than one tile is refused with MMBA_ERR_TRACK
    with pytest.raises(_capi.MmbaError) as e:
        _capi.plan(300, 1, np.arange(300, dtype=np.int64), np.zeros(300, dtype=np.int64))
    assert e.value.code == -5


def test_plan_fuzz_ragged_duplicates_and_full_tiles():
    """Seeded random problems across the awkward shapes: duplicate (camera, point) pairs, unobserved points and
    cameras, tracks of exactly one tile (256 observations, the documented limit), single-camera problems, more ranks
    than tiles, unsorted observation order."""
    rng = np.random.default_rng(11)
    for case in range(40):
        nc = int(rng.integers(1, 40))
        npts = int(rng.integers(1, 120))
        lens = rng.integers(0, 12, size=npts)
        if case % 5 == 0:
            lens[rng.integers(npts)] = 256          # a track that fills a tile exactly
        if case % 7 == 0:
            lens[rng.integers(npts)] = 255
        if lens.sum() == 0:
            lens[0] = 1
        pi = np.repeat(np.arange(npts, dtype=np.int64), lens)
        fi = rng.integers(0, nc, size=len(pi)).astype(np.int64)      # duplicates of (camera, point) are allowed
        order = rng.permutation(len(pi))
        _check_plan(nc, npts, fi[order], pi[order], int(rng.integers(1, 5)))


# ---- block pattern of the reduced camera matrix ---------------------------------------------------
@pytest.mark.parametrize("windowed", [True, False])
def test_rcm_pattern_is_the_covisibility_graph(windowed):
    import scipy.sparse as sp
    prob = synth.make_problem(40, 900, 5400, seed=3, hard=False, windowed=windowed)
    ext, K, pts, uv, fi, pi = prob.args()
    nc, npts = len(ext), len(pts)
    rowptr, cols, nnz_full, pairs = _capi.host_rcm_pattern(nc, npts, fi, pi)
    vis = sp.csr_matrix((np.ones(len(fi)), (fi, pi)), shape=(nc, npts))
    cov = ((vis @ vis.T).toarray() > 0) | np.eye(nc, dtype=bool)
    dense = np.zeros((nc, nc), dtype=bool)
    for i in range(nc):
        c = cols[rowptr[i]:rowptr[i + 1]]
        assert c[0] == i and np.all(np.diff(c) > 0)
        dense[i, c] = True
    assert np.array_equal(dense, np.triu(cov))
    assert nnz_full == cov.sum()
    L = np.bincount(pi, minlength=npts)
    assert pairs == int((L * (L + 1) // 2).sum())


def test_rcm_pattern_fuzz_against_the_covisibility_graph():
    """Seeded random problems (duplicates, unobserved cameras and points, single-camera and single-point shapes):
    the block pattern is the upper triangle of the co-visibility graph plus every diagonal block."""
    import scipy.sparse as sp
    rng = np.random.default_rng(5)
    for _ in range(40):
        nc, npts = int(rng.integers(1, 30)), int(rng.integers(1, 80))
        nobs = int(rng.integers(1, 400))
        fi = rng.integers(0, nc, size=nobs).astype(np.int64)
        pi = rng.integers(0, npts, size=nobs).astype(np.int64)
        rowptr, cols, nnz_full, pairs = _capi.host_rcm_pattern(nc, npts, fi, pi)
        vis = sp.csr_matrix((np.ones(nobs), (fi, pi)), shape=(nc, npts))
        cov = ((vis @ vis.T).toarray() > 0) | np.eye(nc, dtype=bool)
        dense = np.zeros((nc, nc), dtype=bool)
        for i in range(nc):
            c = cols[rowptr[i]:rowptr[i + 1]]
            assert c[0] == i and np.all(np.diff(c) > 0)
            dense[i, c] = True
        assert np.array_equal(dense, np.triu(cov)) and nnz_full == cov.sum()
        L = np.bincount(pi, minlength=npts)
        assert pairs == int((L * (L + 1) // 2).sum())


def test_rcm_pattern_unobserved_camera_keeps_its_diagonal():
    fi = np.array([0, 2, 0, 2])
    pi = np.array([0, 0, 1, 1])
    rowptr, cols, nnz_full, pairs = _capi.host_rcm_pattern(4, 2, fi, pi)
    assert rowptr.tolist() == [0, 2, 3, 4, 5] and cols.tolist() == [0, 2, 1, 2, 3]
    assert nnz_full == 6 and pairs == 6


# ---- tile strategies of the S-build pass / ring-aware point order --------------------------------
def _tile_stats(prob):
    ext, K, pts, uv, fi, pi = prob.args()
    fi = np.ascontiguousarray(fi, np.int64)
    pi = np.ascontiguousarray(pi, np.int64)
    h = _capi._H()
    _capi._check(_capi.lib().mmba_plan_create(ctypes.byref(h), len(ext), len(pts), len(fi), fi, pi, 0, 1))
    sizes = (ctypes.c_int64 * 8)()
    _capi.lib().mmba_plan_sizes(h, ctypes.byref(sizes))
    n = int(sizes[0])
    out = [np.zeros(n, dtype=np.int32) for _ in range(4)]
    _capi._check(_capi.lib().mmba_plan_tile_stats(h, *out))
    _capi.lib().mmba_plan_destroy(h)
    return out


def test_turntable_tiles_stay_narrow_and_register_resident():
    """Tracks that wrap around the camera ring are keyed by their first camera in the upper half: every tile of the
    video-like shape sees about one window of cameras and takes the register-resident S-build strategy."""
    prob = synth.make_config("C2", hard=True, scale=0.05)          # 200 cameras, windows of 20, ring closed
    ncams, npts, mode, npairs = _tile_stats(prob)
    assert ncams.max() <= 24 and np.all(mode == 2)
    L = np.bincount(prob.pt_idx)
    assert npairs.sum() == int((L * (L + 1) // 2).sum())


def test_random_visibility_and_duplicates_take_the_general_strategy():
    prob = synth.make_problem(300, 900, 4000, seed=9, windowed=False)
    ncams, npts, mode, npairs = _tile_stats(prob)
    assert ncams.max() > 100 and np.all(mode == 0)
    dup = synth.make_problem(12, 300, 1500, seed=31)
    dup.cam_idx = np.concatenate((dup.cam_idx, dup.cam_idx[:50]))
    dup.pt_idx = np.concatenate((dup.pt_idx, dup.pt_idx[:50]))
    dup.uv = np.vstack((dup.uv, dup.uv[:50]))
    _, _, mode, _ = _tile_stats(dup)
    assert (mode == 0).any()                                       # tiles holding a duplicated observation


def test_rcm_pattern_does_not_depend_on_the_observation_order():
    prob = synth.make_problem(50, 1200, 6000, seed=4)
    ext, K, pts, uv, fi, pi = prob.args()
    a = _capi.host_rcm_pattern(len(ext), len(pts), fi, pi)
    perm = np.random.default_rng(1).permutation(len(fi))
    b = _capi.host_rcm_pattern(len(ext), len(pts), fi[perm], pi[perm])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2:] == b[2:]
