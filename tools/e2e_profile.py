"""Where an adjustPoints call spends its time (host + device phases), per config.
    MMBA_PLAN_TIMING=1 python tools/e2e_profile.py [C2] [C4]"""
import contextlib, io, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meatmodeler_b200 import synth
from meatmodeler_b200 import bundleAdjuster as mm

for name in sys.argv[1:] or ["C2"]:
    prob = synth.make_config(name, hard=True)
    ext, K, pts, uv, fi, pi = prob.args()
    nc, npts, nobs = prob.sizes
    sink = io.StringIO()
    for rep in range(3):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(sink):
            mm.adjustPoints(ext, K, pts, uv, fi, pi)
        t1 = time.perf_counter()
        print(f"{name} adjustPoints call {rep}: {1e3 * (t1 - t0):.2f} ms (solve {mm.last_result.solve_ms:.2f} ms, nit {mm.last_result.nit}, "
              f"pcg {mm.last_result.pcg_iterations}, cost {mm.last_result.cost:.6f})", flush=True)
    # phases of the Python wrapper
    t = time.perf_counter()
    x0 = np.hstack((mm.frameParameters(np.asarray(ext, dtype=np.float64)), np.asarray(pts, dtype=np.float64).reshape(-1)))
    t_pack = time.perf_counter() - t
    eng = next(iter(mm._ENGINES.values()))
    t = time.perf_counter(); eng.set_problem(nc, npts, K, fi, pi, uv); t_set = time.perf_counter() - t
    t = time.perf_counter(); x, r, _ = eng.solve(x0); t_solve = time.perf_counter() - t
    t = time.perf_counter(); out = mm.reformatPointResult(mm.SolveResult(x=x), nc, npts); t_unpack = time.perf_counter() - t
    print(f"{name} phases: pack {1e3 * t_pack:.2f} ms | set_problem {1e3 * t_set:.2f} ms | Engine.solve {1e3 * t_solve:.2f} ms "
          f"(device {r.solve_ms:.2f}) | unpack {1e3 * t_unpack:.2f} ms", flush=True)
