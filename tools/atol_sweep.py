"""pcg_atol sweep on the full-size configs against the reference's golden trajectories (tests/golden/c2.npz, c4.npz)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meatmodeler_b200 import synth
from meatmodeler_b200 import bundleAdjuster as mm

for name in sys.argv[1:] or ["C4"]:
    prob = synth.make_config(name, hard=True)
    ext, K, pts, uv, fi, pi = prob.args()
    x0 = np.hstack((mm.frameParameters(ext), np.asarray(pts).reshape(-1)))
    g = np.load(os.path.join(ROOT, "tests", "golden", name.lower() + ".npz"))
    print(name, "reference: nfev", int(g["ref_nfev"]), "lsmr", g["ref_lsmr_its"], "costs", g["ref_costs"])
    for atol in (1e-7, 3e-7, 1e-6, 3e-6, 1e-5):
        res = mm.solve(x0, K, len(ext), len(pts), fi, pi, uv, pcg_atol=atol)
        costs = np.array([r["cost"] for r in res.log])
        its = [r["pcg_iterations"] for r in res.log][:-1]
        n = min(len(costs), len(g["ref_costs"]))
        print(f"{name} pcg_atol={atol:.0e} nfev={res.nfev} status={res.status} pcg={its} final={res.cost:.6f} "
              f"rel_vs_ref={(res.cost - float(g['ref_cost'])) / float(g['ref_cost']):+.2e} solve_ms={res.solve_ms:.1f} "
              f"max_rel_traj={np.max(np.abs(costs[:n] - g['ref_costs'][:n]) / g['ref_costs'][:n]):.1e}")
