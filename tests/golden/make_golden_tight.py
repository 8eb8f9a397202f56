"""Golden trajectories of the UNMODIFIED reference on long camera chains with its inner solves CONVERGED.

The reference's call (bundleAdjuster.py:180-192) leaves LSMR at scipy's default atol = btol = 1e-6; on chain-like
camera graphs (video-like visibility, hundreds of cameras) LSMR then stops far from the solution of the damped
Gauss-Newton system and the reference's own final cost is only defined to ~1e-4 (profiles/r2_ref_lsmr_sensitivity_*).
Here the same call runs with ``tr_options=dict(atol=tol, btol=tol)``, tol = 1e-11 — the reference's residual /
sparsity / packing code and scipy's TRF untouched, only the inner tolerance tightened — which gives a trajectory that
IS well defined (tol = 1e-9 and 1e-11 agree to 7e-9 on the final cost): the one an exact inner solver must reproduce to
the north-star bar.  The default-tolerance trajectory of the same problem rides along.

Run in the build container only (needs /root/reference), single-threaded BLAS (the sparse products are serial anyway):
    OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 python tests/golden/make_golden_tight.py [chain] [c4s]
  chain: 300 cameras on the ring, 6 000 points, 30 000 observations (windows of 5 neighbours) — about 2 minutes
  chain1k: 1 000 cameras, 20 000 points, 100 000 observations (windows of 5 neighbours), LSMR at 1e-10
  c4s  : BASELINE configs[3] scaled to 5 % of its points, all 1 778 cameras kept (250 k observations) — hours
Writes tests/golden/<name>_tight.npz.  Nothing at test time reads /root/reference.
"""
import os
import sys
import time

import numpy as np
import scipy
from scipy.optimize import least_squares

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402  (imports the reference by path, unmodified)
from meatmodeler_b200 import synth  # noqa: E402

ref = mg.ref
TOL = 1e-11

PROBLEMS = {
    "chain": lambda: synth.make_problem(300, 6000, 30000, seed=33, hard=True),
    "chain1k": lambda: synth.make_problem(1000, 20000, 100000, seed=34, hard=True),
    "c4s": lambda: synth.make_config("C4", hard=True, scale=0.05),
}
TOLS = {"chain1k": 1e-10, "c4s": 1e-10}


def trajectory(prob, tol):
    """The reference call with LSMR's tolerances set to ``tol`` (None: the reference's own call, scipy's 1e-6)."""
    import scipy.optimize._lsq.trf as T
    ext, K, pts, uv, fi, pi = prob.args()
    nc, npts = len(ext), len(pts)
    x0 = np.hstack((ref.frameParameters(ext), pts.reshape(npts * 3)))
    A = ref.pointAdjustmentSparsity(nc, npts, fi, pi)
    costs, its = [], []
    orig = T.lsmr

    def spy(*a, **k):
        out = orig(*a, **k)
        its.append(int(out[2]))
        return out

    T.lsmr = spy
    try:
        kw = {} if tol is None else dict(tr_options=dict(atol=tol, btol=tol))
        res = least_squares(ref.pointFun, x0, jac_sparsity=A, verbose=0, x_scale="jac", ftol=1e-4, method="trf",
                            args=(K, nc, npts, fi, pi, uv),
                            callback=lambda intermediate_result: costs.append(float(intermediate_result.cost)), **kw)
    finally:
        T.lsmr = orig
    f0 = ref.pointFun(x0, K, nc, npts, fi, pi, uv)
    rms = np.sqrt(np.mean(np.sum(res.fun.reshape(-1, 2) ** 2, axis=1)))
    return x0, res, np.array([0.5 * f0 @ f0] + costs), np.array(its), rms


def main(names):
    for name in names:
        prob = PROBLEMS[name]()
        uv = prob.args()[3]
        t0 = time.perf_counter()
        tol = TOLS.get(name, TOL)
        x0, res, costs, its, rms = trajectory(prob, tol)
        wall = time.perf_counter() - t0
        _, res_d, costs_d, its_d, rms_d = trajectory(prob, None)
        np.savez_compressed(
            os.path.join(HERE, name + "_tight.npz"), sizes=np.array(prob.sizes), x0_checksum=float(np.sum(x0)),
            uv_checksum=float(np.sum(uv)), lsmr_tol=tol, ref_costs=costs, ref_cost=res.cost, ref_nfev=res.nfev,
            ref_status=res.status, ref_rms=rms, ref_lsmr_its=its, ref_wall_s=wall,
            ref_default_costs=costs_d, ref_default_cost=res_d.cost, ref_default_nfev=res_d.nfev,
            ref_default_status=res_d.status, ref_default_rms=rms_d, ref_default_lsmr_its=its_d,
            versions=np.array([np.__version__, scipy.__version__]))
        print(name, prob.sizes, f"tol {tol:g}: nfev {res.nfev} status {res.status} cost {res.cost:.9f} lsmr {its.tolist()} "
              f"{wall:.0f} s | default: nfev {res_d.nfev} cost {res_d.cost:.9f} lsmr {its_d.tolist()}", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["chain"])
