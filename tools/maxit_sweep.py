import sys
import numpy as np
sys.path.insert(0, "/root/repo")
from meatmodeler_b200 import synth
from meatmodeler_b200 import bundleAdjuster as mm
prob = synth.make_config("C4", hard=True)
ext, K, pts, uv, fi, pi = prob.args()
x0 = np.hstack((mm.frameParameters(ext), np.asarray(pts).reshape(-1)))
g = np.load("/root/repo/tests/golden/c4.npz")
for maxit in (500, 300, 200):
    res = mm.solve(x0, K, len(ext), len(pts), fi, pi, uv, pcg_maxit=maxit)
    its = [r["pcg_iterations"] for r in res.log][:-1]
    costs = np.array([r["cost"] for r in res.log])
    print(f"maxit={maxit} nfev={res.nfev} status={res.status} pcg={its} final={res.cost:.4f} rel_vs_ref={(res.cost - float(g['ref_cost'])) / float(g['ref_cost']):+.2e} solve_ms={res.solve_ms:.1f} costs={np.array2string(costs, precision=1)}")
