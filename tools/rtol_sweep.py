"""PCG tolerance sweep: iteration counts and cost trajectories vs the reference golden files."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meatmodeler_b200 import synth
from meatmodeler_b200 import bundleAdjuster as mm

def x0_of(prob):
    ext, K, pts, uv, fi, pi = prob.args()
    return np.hstack((mm.frameParameters(ext), np.asarray(pts).reshape(-1)))

cases = {"c1": synth.make_config("C1", hard=True),
         "mid": synth.make_problem(60, 1500, 9000, seed=7, hard=True, windowed=False),
         "C2": synth.make_config("C2", hard=True),
         "C4": synth.make_config("C4", hard=True)}
for name, prob in cases.items():
    ext, K, pts, uv, fi, pi = prob.args()
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")) if name in ("c1", "mid") else None
    base = None
    for rtol in (1e-10, 1e-8, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2):
        res = mm.solve(x0_of(prob), K, len(ext), len(pts), fi, pi, uv, pcg_rtol=rtol)
        costs = np.array([r["cost"] for r in res.log])
        its = [r["pcg_iterations"] for r in res.log][:-1]
        if base is None:
            base = costs
        ref = g["ref_costs"] if g is not None else base
        n = min(len(ref), len(costs))
        rel = np.abs(costs[:n] - ref[:n]) / ref[:n]
        print(f"{name} rtol={rtol:.0e} nfev={res.nfev} nit={res.nit} pcg={its} final={res.cost:.10e} "
              f"solve_ms={res.solve_ms:.2f} max_rel_vs_{'ref' if g is not None else '1e-10'}={rel.max():.2e} final_rel={rel[-1]:.2e} len={len(costs)}/{len(ref)}")
