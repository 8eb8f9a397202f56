"""The explicit reduced camera matrix (schur_mode = EXPLICIT): S-build pass + one-kernel PCG against a numpy
Schur complement built from the oracle's Jacobian blocks, against the oracle's PCG / TRF, and against the
implicit path of the same engine."""
import numpy as np
import pytest
import scipy.sparse as sp

from meatmodeler_b200 import _capi, synth
from oracle import schur_trf

from conftest import problem_x0
from test_gpu_parity import PROBLEMS, engine_for

pytestmark = pytest.mark.gpu


def _with_duplicates():
    """A camera observing the same point twice (two slightly different pixels): legal input for the reference."""
    prob = synth.make_problem(12, 300, 1500, seed=31, hard=True)
    rng = np.random.default_rng(4)
    pick = rng.choice(len(prob.uv), 200, replace=False)
    prob.uv = np.vstack((prob.uv, prob.uv[pick] + rng.normal(0, 0.3, (200, 2))))
    prob.cam_idx = np.concatenate((prob.cam_idx, prob.cam_idx[pick]))
    prob.pt_idx = np.concatenate((prob.pt_idx, prob.pt_idx[pick]))
    return prob


CASES = dict(PROBLEMS)
CASES["duplicates"] = _with_duplicates
CASES["long_tracks"] = lambda: synth.make_problem(150, 60, 6000, seed=5, hard=True)   # 100 observations per point
# 2 observations per point: 128 points per tile (the producer's gather share is exceeded, the consumers complete it)
CASES["short_tracks"] = lambda: synth.make_problem(40, 6000, 12000, seed=7, hard=True)


@pytest.fixture(scope="module", params=sorted(CASES))
def xcase(request):
    prob = CASES[request.param]()
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(prob.uv))
    prob.uv, prob.cam_idx, prob.pt_idx = prob.uv[perm], prob.cam_idx[perm], prob.pt_idx[perm]
    x0 = problem_x0(prob)
    ext, K, pts, uv, fi, pi = prob.args()
    lin = schur_trf.Linearisation(x0, K, len(ext), len(pts), fi, pi, uv)
    eng = engine_for(prob, pcg_rtol=1e-10, pcg_atol=0.0, pcg_ktol=0.0, schur_mode=_capi.SCHUR_EXPLICIT)
    yield prob, x0, lin, eng
    eng.close()


def numpy_reduced_system(lin, d, reg):
    """S = D_c (U - W V'^-1 W^T) D_c + reg I and b = D_c (g_c - W V'^-1 g_p), dense, from the oracle's blocks."""
    no, nc, npt = len(lin.fi), lin.Nc, lin.Np
    rows = np.repeat(np.arange(2 * no), 6)
    Jc = sp.csr_matrix((lin.Jc.reshape(-1), (rows, (6 * lin.fi[:, None, None] + np.arange(6)[None, None, :] +
                                                    np.zeros((1, 2, 1), dtype=np.int64)).reshape(-1))), shape=(2 * no, 6 * nc))
    rows3 = np.repeat(np.arange(2 * no), 3)
    Jp = sp.csr_matrix((lin.Jp.reshape(-1), (rows3, (3 * lin.pi[:, None, None] + np.arange(3)[None, None, :] +
                                                     np.zeros((1, 2, 1), dtype=np.int64)).reshape(-1))), shape=(2 * no, 3 * npt))
    dc, dp = d[:6 * nc], d[6 * nc:]
    Jc = Jc @ sp.diags(dc)
    Jp = Jp @ sp.diags(dp)
    U = (Jc.T @ Jc).toarray()
    W = (Jc.T @ Jp).tocsr()
    V = (Jp.T @ Jp + reg * sp.identity(3 * npt)).tocsc()
    Vinv = sp.linalg.inv(V) if npt * 3 < 4000 else None
    g = d * lin.grad()
    if Vinv is None:
        lu = sp.linalg.splu(V)
        WVi = lu.solve(W.T.toarray()).T
    else:
        WVi = (W @ Vinv).toarray()
    S = U + reg * np.eye(6 * nc) - WVi @ W.T.toarray()
    b = g[:6 * nc] - WVi @ g[6 * nc:]
    return S, b


@pytest.mark.parametrize("reg", [1e-2, 1e-6])
def test_reduced_system_vs_numpy(xcase, reg):
    prob, x0, lin, eng = xcase
    d = 1.0 / np.where(lin.colnorm() == 0, 1.0, lin.colnorm())
    S, b = eng.reduced_system(x0, d, reg)
    S_ref, b_ref = numpy_reduced_system(lin, d, reg)
    assert np.abs(S - S.T).max() <= 1e-12 * np.abs(S).max()
    assert np.abs(S - S_ref).max() <= 1e-10 * np.abs(S_ref).max()
    assert np.abs(b - b_ref).max() <= 1e-10 * np.abs(b_ref).max()


@pytest.mark.parametrize("reg", [1e-2, 1e-6])
def test_explicit_gauss_newton_step_vs_oracle(xcase, reg):
    prob, x0, lin, eng = xcase
    d = 1.0 / np.where(lin.colnorm() == 0, 1.0, lin.colnorm())
    p, its, rel = eng.gn_step(x0, d, reg)
    p_ref, its_ref, rel_ref = schur_trf.schur_pcg(lin, d, reg, 1e-10, 1000)
    # Same method and preconditioner on both sides.  The engine's one-exchange-per-iteration form carries z = Pinv r
    # and the next inner products by recurrence; on the ill-conditioned system (reg = 1e-6) pushed to rtol = 1e-10 that
    # costs up to a third more iterations than the textbook recurrences, varying with the summation order of the
    # S-build's atomics (measured 97 .. 130 against the oracle's 97).  The default stopping rules end far earlier.
    assert rel <= 1e-10 and its <= its_ref + max(3, its_ref // 2), (rel, its, its_ref)
    Jdp = lin.jdot(d * p)
    nc = lin.Nc
    JtJdp = np.hstack((schur_trf._segsum(lin.fi, np.einsum("nij,ni->nj", lin.Jc, Jdp), nc).ravel(),
                       schur_trf._segsum(lin.pi, np.einsum("nij,ni->nj", lin.Jp, Jdp), lin.Np).ravel()))
    rhs = d * lin.grad()
    resid = d * JtJdp + reg * p - rhs
    assert np.linalg.norm(resid) <= 1e-8 * np.linalg.norm(rhs)
    # The step itself is only determined to cond(S) x the residual: at reg = 1e-6 the 'windowed' chain has near-gauge
    # directions, and two solves that both reach 1e-10 relative residual differ by up to 1.3e-5 in p depending on the
    # summation order of the S-build's atomics (tools/pcg_flaky.py: 6 of 8 runs 1e-9, 2 of 8 runs 1.29e-5).  The
    # residual bar above is the parity statement; this one only guards against a wrong solution.
    assert np.linalg.norm(p - p_ref) <= (1e-5 if reg >= 1e-3 else 1e-4) * np.linalg.norm(p_ref)


def test_explicit_solve_vs_oracle_and_implicit(xcase):
    prob, x0, lin, eng = xcase
    ext, K, pts, uv, fi, pi = prob.args()
    rec = []
    out = schur_trf.solve(x0, K, len(ext), len(pts), fi, pi, uv, record=rec, pcg_atol=0.0, pcg_ktol=0.0)
    x, r, fun = eng.solve(x0, want_fun=True)
    costs = [row["cost"] for row in eng.log()][1:]
    assert r.nfev == out["nfev"] and r.status == out["status"] and len(costs) == len(rec)
    # two observations per point ('short_tracks') is weakly determined geometry: the 1e-10 inner solves of the two
    # implementations differ by cond(S) x 1e-10 in the step, 1.7e-7 in the intermediate costs (measured); the reduced
    # system itself and the Gauss-Newton step of that case pass the 1e-10 / 1e-8 bars above
    weak = len(prob.points) == 6000
    np.testing.assert_allclose(costs, rec, rtol=1e-6 if weak else 1e-8)
    assert r.cost == pytest.approx(out["cost"], rel=1e-7 if weak else 1e-9)
    # the same engine with the implicit product (the mode can be lowered on a live handle)
    eng.set_options(schur_mode=_capi.SCHUR_IMPLICIT)
    x2, r2, _ = eng.solve(x0)
    eng.set_options(schur_mode=_capi.SCHUR_EXPLICIT)
    assert r2.nfev == r.nfev and r2.status == r.status
    assert r2.cost == pytest.approx(r.cost, rel=1e-7 if weak else 1e-9)
    assert r.pcg_iterations > 0


def test_auto_mode_picks_explicit_for_video_like_visibility_only():
    dense = synth.make_problem(300, 900, 4000, seed=9, windowed=False)
    video = synth.make_config("C1", hard=True)
    x0 = problem_x0(video)
    with engine_for(video) as eng:
        d = np.ones(x0.size)
        eng.reduced_system(x0, d, 1e-3)                     # formed: no error
    with engine_for(dense) as eng:
        with pytest.raises(_capi.MmbaError) as e:
            eng.reduced_system(problem_x0(dense), np.ones(problem_x0(dense).size), 1e-3)
        assert e.value.code == -3


def test_lost_pcg_cta_is_an_error_code_not_a_hang(monkeypatch):
    """The one-kernel PCG exchanges through spin-polled lines; a CTA that never publishes (fault injection) makes
    the others give up after their spin limit (~1 s): the solve returns MMBA_ERR_CUDA, and the handle stays usable."""
    prob = synth.make_problem(40, 3000, 24000, seed=21, hard=True)
    x0 = problem_x0(prob)
    ext, K, pts, uv, fi, pi = prob.args()
    with _capi.Engine(schur_mode=_capi.SCHUR_EXPLICIT) as eng:
        eng.set_problem(len(ext), len(pts), K, fi, pi, uv)
        if eng.rcm_pattern()["n_ctas"] < 2:
            pytest.skip("needs at least two PCG CTAs")
        monkeypatch.setenv("MMBA_FAULT_PCG_CTA", "1")
        with pytest.raises(_capi.MmbaError, match="timed out"):
            eng.solve(x0)
        monkeypatch.delenv("MMBA_FAULT_PCG_CTA")
        x, r, _ = eng.solve(x0)
        assert r.status > 0 and np.isfinite(r.cost)


def test_reduced_system_with_tiles_of_single_observation_points():
    """700 points seen once each by two cameras: tiles of 256 points, whose payloads no longer fit a three-stage
    S-build pipeline (the kernel drops to two stages); the reduced system must still equal the numpy Schur complement."""
    prob = synth.make_problem(30, 500, 3000, seed=11, hard=True)
    rng = np.random.default_rng(5)
    extra = 700
    src = rng.integers(0, len(prob.points), extra)
    cams = np.repeat([3, 17], extra // 2)
    pts_true = prob.true_points[src] + rng.normal(0, 0.05, (extra, 3))
    uv_extra = synth._project(prob.true_extrinsics, prob.K, pts_true, cams, np.arange(extra)) + rng.normal(0, 0.5, (extra, 2))
    prob.points = np.concatenate((prob.points, (pts_true + rng.normal(0, 0.02, (extra, 3)))[:, None, :]))
    prob.true_points = np.concatenate((prob.true_points, pts_true))
    prob.uv = np.vstack((prob.uv, uv_extra))
    prob.cam_idx = np.concatenate((prob.cam_idx, cams))
    prob.pt_idx = np.concatenate((prob.pt_idx, 500 + np.arange(extra)))
    x0 = problem_x0(prob)
    ext, K, pts, uv, fi, pi = prob.args()
    lin = schur_trf.Linearisation(x0, K, len(ext), len(pts), fi, pi, uv)
    d = 1.0 / np.where(lin.colnorm() == 0, 1.0, lin.colnorm())
    with engine_for(prob, schur_mode=_capi.SCHUR_EXPLICIT) as eng:
        tile_points = np.ascontiguousarray(eng.plan()["meta"][:, 4:8]).view(np.int32).reshape(-1)    # TileMeta.npts
        assert tile_points.max() >= 200
        S, b = eng.reduced_system(x0, d, 1e-2)
    S_ref, b_ref = numpy_reduced_system(lin, d, 1e-2)
    assert np.abs(S - S_ref).max() <= 1e-10 * np.abs(S_ref).max()
    assert np.abs(b - b_ref).max() <= 1e-10 * np.abs(b_ref).max()
