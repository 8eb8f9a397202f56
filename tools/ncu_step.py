"""One solve of one config between cudaProfilerStart / Stop (for `ncu --profile-from-start off`).
    python tools/ncu_step.py [C4]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from meatmodeler_b200 import _capi, synth
from meatmodeler_b200 import bundleAdjuster as mm

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
prob = synth.make_config(name, hard=True)
ext, K, pts, uv, fi, pi = prob.args()
x0 = np.hstack((mm.frameParameters(ext), np.asarray(pts).reshape(-1)))
with _capi.Engine() as eng:
    eng.set_problem(len(ext), len(pts), K, fi, pi, uv)
    eng.set_x(x0)
    for _ in range(3):
        eng.solve_resident()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    r = eng.solve_resident()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print(f"{name}: solve {r.solve_ms:.3f} ms nit {r.nit} nfev {r.nfev} pcg {r.pcg_iterations} cost {r.cost:.6f}")
