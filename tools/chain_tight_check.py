"""Engine vs the unmodified reference with CONVERGED inner solves on long camera chains (prints, no assertions).
    python tools/chain_tight_check.py [--quick] [chain] [chain1k] [c4s]      (--quick: default rules + one converged setting)
Goldens: tests/golden/<name>_tight.npz (make_golden_tight.py: LSMR at 1e-11 instead of scipy's 1e-6).  For every
setting of the engine's inner solve: nfev / status, final cost and its relative deviation from the converged reference,
largest relative deviation along the trajectory, PCG iterations per inner solve, device time."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from meatmodeler_b200 import _capi, synth
from meatmodeler_b200 import bundleAdjuster as mm

PROBLEMS = {
    "chain": lambda: synth.make_problem(300, 6000, 30000, seed=33, hard=True),
    "chain1k": lambda: synth.make_problem(1000, 20000, 100000, seed=34, hard=True),
    "c4s": lambda: synth.make_config("C4", hard=True, scale=0.05),
}
TIGHT = dict(pcg_atol=0.0, pcg_ktol=0.0, pcg_maxit=100000)
SETTINGS = [
    ("default rules", {}),
    ("converged, rtol 1e-8", dict(TIGHT, pcg_rtol=1e-8)),
    ("converged, rtol 1e-9", dict(TIGHT, pcg_rtol=1e-9)),
    ("converged, rtol 1e-10", dict(TIGHT, pcg_rtol=1e-10)),
    ("converged, rtol 1e-9, implicit Schur product", dict(TIGHT, pcg_rtol=1e-9, schur_mode=_capi.SCHUR_IMPLICIT)),
]

args = [a for a in sys.argv[1:] if not a.startswith("--")]
if "--quick" in sys.argv:
    SETTINGS = [SETTINGS[0], ("converged, rtol 1e-9", dict(TIGHT, pcg_rtol=1e-9, pcg_maxit=400000))]
for name in args or ["chain"]:
    path = os.path.join(ROOT, "tests", "golden", name + "_tight.npz")
    if not os.path.exists(path):
        print(name, "no golden", path)
        continue
    g = np.load(path)
    prob = PROBLEMS[name]()
    ext, K, pts, uv, fi, pi = prob.args()
    x0 = np.hstack((mm.frameParameters(ext), np.asarray(pts).reshape(-1)))
    assert abs(x0.sum() - float(g["x0_checksum"])) < 1e-9
    ref, ref_cost, ref_rms = g["ref_costs"], float(g["ref_cost"]), float(g["ref_rms"])
    print(f"{name} {prob.sizes}: reference converged (LSMR {float(g['lsmr_tol']):g}): nfev {int(g['ref_nfev'])} status {int(g['ref_status'])} "
          f"cost {ref_cost:.9f} lsmr {g['ref_lsmr_its'].tolist()} | reference default: nfev {int(g['ref_default_nfev'])} "
          f"cost {float(g['ref_default_cost']):.9f} ({(float(g['ref_default_cost']) - ref_cost) / ref_cost:+.2e}) "
          f"lsmr {g['ref_default_lsmr_its'].tolist()}", flush=True)
    for label, opts in SETTINGS:
        try:
            res = mm.solve(x0, K, len(ext), len(pts), fi, pi, uv, want_fun=True, **opts)
        except Exception as e:      # noqa: BLE001 - diagnostics: report and go on
            print(f"  {label}: FAILED {type(e).__name__}: {e}", flush=True)
            continue
        costs = np.array([r["cost"] for r in res.log])
        n = min(len(costs), len(ref))
        rms = np.sqrt(np.mean(np.sum(res.fun.reshape(-1, 2) ** 2, axis=1)))
        print(f"  {label}: nfev {res.nfev} status {res.status} cost {res.cost:.9f} rel {(res.cost - ref_cost) / ref_cost:+.2e} "
              f"rms rel {(rms - ref_rms) / ref_rms:+.2e} traj {np.max(np.abs(costs[:n] - ref[:n]) / ref[:n]):.2e} "
              f"pcg {[int(r['pcg_iterations']) for r in res.log]} solve {res.solve_ms:.2f} ms "
              f"costs {[f'{c:.6f}' for c in costs]}", flush=True)
