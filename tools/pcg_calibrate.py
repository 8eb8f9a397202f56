"""GPU: residual history of the reduced-system PCG on the full-size configs and a sweep of the LSMR-like
k-dependent stop (pcg_ktol) against the reference's golden trajectories.  Writes gpurun_out/pcg_hist_<cfg>.npz."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meatmodeler_b200 import _capi, synth
from meatmodeler_b200 import bundleAdjuster as mm

out_dir = os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)
tag = "single"
for name in sys.argv[1:] or ["C4", "C2"]:
    prob = synth.make_config(name, hard=True)
    ext, K, pts, uv, fi, pi = prob.args()
    x0 = np.hstack((mm.frameParameters(ext), np.asarray(pts).reshape(-1)))
    gpath = os.path.join(ROOT, "tests", "golden", name.lower() + ".npz")
    g = np.load(gpath) if os.path.exists(gpath) else None
    if g is not None:
        print(name, "reference: nfev", int(g["ref_nfev"]), "lsmr", g["ref_lsmr_its"], "costs", g["ref_costs"], flush=True)
    eng = _capi.Engine(profile=2)
    eng.set_problem(len(ext), len(pts), K, fi, pi, uv)
    eng.set_x(x0)
    sweep = ((1e-7, 0.0), (1e-7, 1.23e-6)) if tag == "classic" else (
        (1e-7, 0.0), (1e-7, 3e-7), (1e-7, 6e-7), (1e-7, 1.23e-6), (1e-7, 2.5e-6), (1e-7, 5e-6), (1e-7, 1e-5), (1e-7, 2e-5))
    for atol, ktol in sweep:
        eng.set_options(pcg_atol=atol, pcg_ktol=ktol)
        eng.solve_resident()
        r = eng.solve_resident()
        log = eng.log()
        costs = np.array([row["cost"] for row in log])
        its = [row["pcg_iterations"] for row in log][:-1]
        line = (f"{name} [{tag}] atol={atol:.0e} ktol={ktol:.2e} nfev={r.nfev} status={r.status} pcg={its} total={r.pcg_iterations} "
                f"final={r.cost:.6f} solve_ms={r.solve_ms:.2f}")
        if g is not None:
            n = min(len(costs), len(g["ref_costs"]))
            line += (f" rel_vs_ref={(r.cost - float(g['ref_cost'])) / float(g['ref_cost']):+.2e} "
                     f"max_rel_traj={np.max(np.abs(costs[:n] - g['ref_costs'][:n]) / g['ref_costs'][:n]):.1e}")
        print(line, flush=True)
        if ktol == 0.0 and tag == "single":
            hist = eng.pcg_history()
            np.savez_compressed(os.path.join(out_dir, f"pcg_hist_{name}.npz"), costs=costs,
                                **{f"h{i}": h for i, h in enumerate(hist)})
            for i, h in enumerate(hist):
                rr = np.sqrt(h[:, 0] / (2 * costs[i]))          # ||r_k|| / ||f||
                nu = 1 / np.sqrt(np.cumsum(1 / np.maximum(h[:, 0], 1e-300))) / np.sqrt(2 * costs[i])
                ks = [k for k in (1, 10, 30, 100, 200, 300, 500, 700, 1000) if k < len(rr)]
                print(f"  outer {i}: its {len(rr) - 1}  |r|/|f| " + " ".join(f"k{k}:{rr[k]:.1e}" for k in ks), flush=True)
                print(f"            nu/|f|/sqrt(k) " + " ".join(f"k{k}:{nu[k] / np.sqrt(k):.1e}" for k in ks), flush=True)
    eng.close()
    # per-iteration cost of the PCG kernel: CUDA events around every launch of one solve
    with _capi.Engine(profile=1) as engp:
        engp.set_problem(len(ext), len(pts), K, fi, pi, uv)
        engp.set_x(x0)
        engp.solve_resident()
        r = engp.solve_resident()
        p = engp.profile()
        print(f"{name} [{tag}] PCG {p['schur_pcg']['ms']:.3f} ms / {r.pcg_iterations} its = "
              f"{1e3 * p['schur_pcg']['ms'] / max(r.pcg_iterations, 1):.2f} us per iteration; solve {r.solve_ms:.2f} ms; "
              + " ".join(f"{k}={v['ms']:.2f}ms/{v['launches']}" for k, v in p.items() if v['launches']), flush=True)
