"""Per-phase cycle counters of the PCG and S-build kernels (options.profile bit 2), one solve per config."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meatmodeler_b200 import _capi, synth
from meatmodeler_b200 import bundleAdjuster as mm

PCG = ["matvec", "sync1", "w+sums", "exchange_total", "update", "loop_top", "  xchg:reduce+publish", "  xchg:poll+sync"]
SB = ["wait_full", "regs+stage", "y_scatter(+barrier)", "tab+sync", "flush", "pairs", "end_sync", "final_flush"]
for name in sys.argv[1:] or ["C4", "C2"]:
    prob = synth.make_config(name, hard=True)
    ext, K, pts, uv, fi, pi = prob.args()
    x0 = np.hstack((mm.frameParameters(ext), np.asarray(pts).reshape(-1)))
    with _capi.Engine(profile=4) as eng:
        eng.set_problem(len(ext), len(pts), K, fi, pi, uv)
        eng.set_x(x0)
        eng.solve_resident()
        eng.phase_cycles(reset=True)
        r = eng.solve_resident()
        c = eng.phase_cycles()
    its = max(r.pcg_iterations, 1)
    print(f"{name}: solve {r.solve_ms:.2f} ms, nit {r.nit}, pcg {r.pcg_iterations}")
    tot = c[0] + c[1] + c[2] + c[3] + c[4] + c[5]
    print(f"  PCG per iteration: {tot / its:.0f} cycles = {tot / its / 1965:.2f} us")
    for k, nm in enumerate(PCG):
        print(f"    {nm:26s} {c[k] / its:8.0f} cycles")
    tiles = max(c[24], 1)
    tot = sum(c[16:24])
    print(f"  S-build CTA 0: {tiles} tile visits ({r.nit} launches), {tot / tiles:.0f} cycles per tile")
    for k, nm in enumerate(SB):
        print(f"    {nm:26s} {c[16 + k] / tiles:8.0f} cycles")
    nfl = max(c[40], 1)
    print(f"  flushes: {c[40]} ({tiles / nfl:.1f} tiles per flush); cycles per flush:")
    for k, nm in enumerate(["rhs reduce", "block index", "stage+barrier", "slice sums+RED", "barrier", "tail"]):
        print(f"    {nm:26s} {c[32 + k] / nfl:8.0f} cycles")
    print(f"    {'block lookup (run start)':26s} {c[41] / nfl:8.0f} cycles")
