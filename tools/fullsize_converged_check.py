"""How far from the CONVERGED answer do the default-tolerance results sit at full size (prints, no assertions)?
    python tools/fullsize_converged_check.py [C3] [C4] [C2]
The unmodified reference cannot be run to convergence at these sizes on a CPU (LSMR would need > 10^5 iterations per
inner solve), but the engine can: with its PCG converged (threshold rules off) it reproduces the converged reference
to 1e-11 on the long-chain golden (tools/chain_tight_check.py).  This script solves BASELINE configs at full size with
the engine's inner solves converged and at its default rules, and compares both, and the reference's own
default-tolerance result (tests/golden/<config>.npz), with the converged cost."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meatmodeler_b200 import synth
from meatmodeler_b200 import bundleAdjuster as mm

for name in sys.argv[1:] or ["C3", "C4"]:
    g = np.load(os.path.join(ROOT, "tests", "golden", name.lower() + ".npz"))
    prob = synth.make_config(name.rstrip("r"), hard=True, windowed=not name.endswith("r"))
    ext, K, pts, uv, fi, pi = prob.args()
    x0 = np.hstack((mm.frameParameters(ext), np.asarray(pts).reshape(-1)))
    assert abs(x0.sum() - float(g["x0_checksum"])) < 1e-6
    ref_cost, ref_nfev = float(g["ref_cost"]), int(g["ref_nfev"])
    runs = {}
    for label, opts in (("default rules", {}),
                        ("converged (rtol 1e-8, threshold rules off)", dict(pcg_rtol=1e-8, pcg_atol=0.0, pcg_ktol=0.0, pcg_maxit=200000)),
                        ("converged (rtol 1e-9, threshold rules off)", dict(pcg_rtol=1e-9, pcg_atol=0.0, pcg_ktol=0.0, pcg_maxit=200000))):
        try:
            runs[label] = mm.solve(x0, K, len(ext), len(pts), fi, pi, uv, **opts)
        except Exception as e:      # noqa: BLE001 - diagnostics: report and go on
            print(f"{name} {label}: FAILED {type(e).__name__}: {e}", flush=True)
    conv = runs.get("converged (rtol 1e-9, threshold rules off)") or runs.get("converged (rtol 1e-8, threshold rules off)")
    c = conv.cost if conv is not None else float("nan")
    print(f"{name} {prob.sizes}: reference (default LSMR tolerance, golden): nfev {ref_nfev} cost {ref_cost:.6f} "
          f"= converged {(ref_cost - c) / c:+.2e}; lsmr {g['ref_lsmr_its'].tolist()}", flush=True)
    for label, res in runs.items():
        costs = ["%.4f" % r["cost"] for r in res.log]
        print(f"  engine, {label}: nfev {res.nfev} status {res.status} cost {res.cost:.6f} = converged {(res.cost - c) / c:+.2e} "
              f"pcg {[int(r['pcg_iterations']) for r in res.log]} solve {res.solve_ms:.2f} ms "
              f"costs {costs}", flush=True)
