"""CPU sweep of the k-dependent PCG stop (oracle/schur_trf.py) against the reference's golden trajectories."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import schur_trf
from meatmodeler_b200 import synth
from conftest import problem_x0

PROBS = {"c1": lambda: synth.make_config("C1", hard=True),
         "mid": None}
for name in sys.argv[1:] or ["c1"]:
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    if name == "small":
        K, fi, pi, uv, x0 = g["K"], g["fi"], g["pi"], g["uv"], g["x0"]; nc, npts = len(g["ext"]), len(g["pts"])
    else:
        import importlib.util
        spec = importlib.util.spec_from_file_location("mg", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
        prob = synth.make_config("C1", hard=True) if name == "c1" else None
        if prob is None:
            raise SystemExit("only c1 / small here")
        ext, K, pts, uv, fi, pi = prob.args(); nc, npts = len(ext), len(pts); x0 = problem_x0(prob)
    ref = g["ref_costs"]
    print(name, "ref lsmr", g["ref_lsmr_its"], "nfev", int(g["ref_nfev"]))
    for atol, ktol in ((1e-7, 0), (0, 1e-7), (0, 3e-7), (0, 6e-7), (0, 1.23e-6), (0, 2.5e-6), (0, 5e-6)):
        rec = []
        out = schur_trf.solve(x0, K, nc, npts, fi, pi, uv, record=rec, pcg_atol=atol, pcg_ktol=ktol)
        costs = np.array([out["log"][0]["cost_before"]] + rec)
        n = min(len(costs), len(ref))
        its = [r["pcg_its"] for r in out["log"]]
        print(f"{name} atol={atol:.0e} ktol={ktol:.2e} nfev={out['nfev']} status={out['status']} pcg={its} final={out['cost']:.10e} "
              f"rel_final={(out['cost']-float(g['ref_cost']))/float(g['ref_cost']):+.1e} max_rel_traj={np.max(np.abs(costs[:n] - ref[:n]) / ref[:n]):.2e} len={len(costs)}/{len(ref)}")
