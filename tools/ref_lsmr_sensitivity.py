"""How well is the reference's OWN final cost determined at C4?  Runs the unmodified reference residual/sparsity code
(build container only: needs /root/reference) through scipy's least_squares with the reference's arguments
(bundleAdjuster.py:180-192) and only LSMR's inner tolerances varied (tr_options atol = btol; scipy's default, the one
the reference gets, is 1e-6), on C4 scaled to `scale` of its points (all 1 778 cameras kept).
    python tools/ref_lsmr_sensitivity.py [scale]      # default 0.05: 250 k observations
The spread of the final costs over the inner tolerance is the accuracy to which "the reference's result" is defined;
a drop-in with ANY other inexact inner solver cannot match it more closely than that.
"""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/reference')
import numpy as np
import bundleAdjuster as ref
from scipy.optimize import least_squares
import scipy.optimize._lsq.trf as T
from meatmodeler_b200 import synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
prob = synth.make_config("C4", hard=True, scale=scale)
ext, K, pts, uv, fi, pi = prob.args()
nc, npts = len(ext), len(pts)
x0 = np.hstack((ref.frameParameters(ext), pts.reshape(npts * 3)))
A = ref.pointAdjustmentSparsity(nc, npts, fi, pi)
orig = T.lsmr
print("sizes", prob.sizes, flush=True)
finals = {}
for tol in (1e-6, 3e-7, 1e-7, 3e-6, 1e-5):
    its = []
    def spy(*a, **k):
        out = orig(*a, **k)
        its.append(int(out[2]))
        return out
    T.lsmr = spy
    costs = []
    t = time.time()
    res = least_squares(ref.pointFun, x0, jac_sparsity=A, verbose=0, x_scale="jac", ftol=1e-4, method="trf",
                        tr_options=dict(atol=tol, btol=tol), args=(K, nc, npts, fi, pi, uv),
                        callback=lambda intermediate_result: costs.append(float(intermediate_result.cost)))
    finals[tol] = res.cost
    print(f"lsmr atol=btol={tol:.0e}{' (reference default)' if tol == 1e-6 else ''}: nfev={res.nfev} status={res.status} "
          f"final cost={res.cost:.6f} rel vs default={(res.cost - finals[1e-6]) / finals[1e-6]:+.2e} lsmr its={its} "
          f"costs={np.array2string(np.array(costs), precision=6)} wall={time.time() - t:.0f}s", flush=True)
T.lsmr = orig
