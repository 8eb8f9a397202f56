"""Deviation of the engine from the reference's golden trajectories at full size (prints, no assertions)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meatmodeler_b200 import synth
from meatmodeler_b200 import bundleAdjuster as mm
for name in sys.argv[1:] or ["C3", "C3r"]:
    g = np.load(os.path.join(ROOT, "tests", "golden", name.lower() + ".npz"))
    prob = synth.make_config(name.rstrip("r"), hard=True, windowed=not name.endswith("r"))
    ext, K, pts, uv, fi, pi = prob.args()
    x0 = np.hstack((mm.frameParameters(ext), np.asarray(pts).reshape(-1)))
    res = mm.solve(x0, K, len(ext), len(pts), fi, pi, uv, want_fun=True)
    costs = np.array([r["cost"] for r in res.log]); ref = g["ref_costs"]; n = min(len(costs), len(ref))
    print(name, "nfev", res.nfev, int(g["ref_nfev"]), "status", res.status, int(g["ref_status"]), "pcg", [r["pcg_iterations"] for r in res.log],
          "lsmr", g["ref_lsmr_its"], "final", res.cost, float(g["ref_cost"]), "rel", (res.cost - float(g["ref_cost"])) / float(g["ref_cost"]),
          "traj", np.max(np.abs(costs[:n] - ref[:n]) / ref[:n]), "solve_ms", res.solve_ms, flush=True)
