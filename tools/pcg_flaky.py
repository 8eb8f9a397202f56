"""Run-to-run spread of the reduced-system PCG pushed to rtol = 1e-10 on the ill-conditioned 'windowed' case, reg = 1e-6."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from meatmodeler_b200 import _capi, synth
from oracle import schur_trf
from conftest import problem_x0
prob = synth.make_problem(40, 3000, 24000, seed=21, hard=True)
rng = np.random.default_rng(0)
perm = rng.permutation(len(prob.uv))
prob.uv, prob.cam_idx, prob.pt_idx = prob.uv[perm], prob.cam_idx[perm], prob.pt_idx[perm]
x0 = problem_x0(prob)
ext, K, pts, uv, fi, pi = prob.args()
lin = schur_trf.Linearisation(x0, K, len(ext), len(pts), fi, pi, uv)
d = 1.0 / np.where(lin.colnorm() == 0, 1.0, lin.colnorm())
for reg in (1e-6,):
    p_ref, its_ref, rel_ref = schur_trf.schur_pcg(lin, d, reg, 1e-10, 1000)
    print("oracle its", its_ref, "rel", rel_ref)
    for rep in range(8):
        with _capi.Engine(pcg_rtol=1e-10, pcg_atol=0.0, pcg_ktol=0.0, schur_mode=_capi.SCHUR_EXPLICIT, profile=2) as eng:
            eng.set_problem(len(ext), len(pts), K, fi, pi, uv)
            p, its, rel = eng.gn_step(x0, d, reg)
            print("rep", rep, "its", its, "rel", rel, "|p - p_ref|/|p_ref|", np.linalg.norm(p - p_ref) / np.linalg.norm(p_ref), flush=True)
            h = eng.pcg_history()
            if h:
                hh = h[-1]
                rr, rho = hh[:, 0], hh[:, 1]
                ratio = rho[1:] / np.maximum(rho[:-1], 1e-300)
                k = np.argsort(ratio)[:3]
                print("   smallest rho ratios at", k, ratio[k], "min rho", rho.min(), "rr tail", np.sqrt(rr[-4:] / rr[0]),
                      "largest rr jump", np.max(rr[1:] / np.maximum(rr[:-1], 1e-300)), flush=True)
