"""Where does the end-to-end adjustPoints call spend its time? (host-side breakdown)"""
import cProfile, io, os, pstats, sys, time, contextlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meatmodeler_b200 import synth, _capi
from meatmodeler_b200 import bundleAdjuster as mm

prob = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "C2", hard=True)
ext, K, pts, uv, fi, pi = prob.args()
sink = io.StringIO()
with contextlib.redirect_stdout(sink):
    mm.adjustPoints(ext, K, pts, uv, fi, pi)
for rep in range(3):
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(sink):
        mm.adjustPoints(ext, K, pts, uv, fi, pi)
    print("adjustPoints wall", round(1e3 * (time.perf_counter() - t0), 2), "ms; solve_ms", round(mm.last_result.solve_ms, 2))
# pieces
nc, npts, nobs = prob.sizes
eng = mm._ENGINES[(0, 0, 1)]
for rep in range(3):
    t0 = time.perf_counter(); eng.set_problem(nc, npts, K, fi, pi, uv); t1 = time.perf_counter()
    x0 = np.hstack((mm.frameParameters(ext), pts.reshape(-1))); t2 = time.perf_counter()
    x, r, _ = eng.solve(x0); t3 = time.perf_counter()
    log = eng.log(); t4 = time.perf_counter()
    class R: pass
    rr = R(); rr.x = x
    mm.reformatPointResult(rr, nc, npts); t5 = time.perf_counter()
    print("set_problem %.2f pack %.2f solve(call) %.2f [device %.2f] log %.2f reformat %.2f ms" % tuple(1e3 * v for v in (t1 - t0, t2 - t1, t3 - t2, r.solve_ms / 1e3, t4 - t3, t5 - t4)))
lib = _capi.lib()
import ctypes as C
for rep in range(3):
    h = _capi._H(); t0 = time.perf_counter()
    lib.mmba_plan_create(C.byref(h), nc, npts, nobs, np.ascontiguousarray(fi), np.ascontiguousarray(pi), 0, 1)
    print("plan_create %.2f ms (OMP threads default; cpu_count %d)" % (1e3 * (time.perf_counter() - t0), os.cpu_count()))
    lib.mmba_plan_destroy(h)
