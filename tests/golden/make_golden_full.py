"""Golden trajectories of the UNMODIFIED reference at BASELINE sizes (configs[1] = C2, configs[2] = C3 / C3r,
configs[3] = C4).

Run in the build container only (needs /root/reference; C2 takes about a minute and 4 GB, C4 about 10 minutes and
18 GB):  python tests/golden/make_golden_full.py [C2] [C4]
Writes tests/golden/c2.npz / c4.npz with the scalars of c1.npz (see make_golden.py): per-iteration costs of the
reference's own ``least_squares`` call, nfev / status, final cost, RMS, LSMR iteration counts, checksums of the
inputs.  Nothing at test time reads /root/reference.
"""
import os
import sys
import time

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402  (imports the reference by path, unmodified)
from meatmodeler_b200 import synth  # noqa: E402


def main(names):
    for name in names:
        # "C3r": BASELINE configs[2] with uniform-random camera subsets per point (the survey's stress variant, dense
        # co-visibility: the engine's implicit Schur path) instead of video-like windows
        random_vis = name.endswith("r")
        prob = synth.make_config(name.rstrip("r"), hard=True, windowed=not random_vis)
        ext, K, pts, uv, fi, pi = prob.args()
        t0 = time.perf_counter()
        x0, res, costs, lsmr_its = mg.reference_trajectory(prob)
        wall = time.perf_counter() - t0
        npts = len(pts)
        rms = np.sqrt(np.mean(np.sum(res.fun.reshape(-1, 2) ** 2, axis=1)))
        np.savez_compressed(
            os.path.join(HERE, name.lower() + ".npz"), sizes=np.array(prob.sizes), x0_checksum=float(np.sum(x0)),
            uv_checksum=float(np.sum(uv)), ref_costs=costs, ref_cost=res.cost, ref_nfev=res.nfev, ref_njev=res.njev,
            ref_status=res.status, ref_optimality=res.optimality, ref_rms=rms, ref_lsmr_its=lsmr_its,
            ref_points_sample=res.x[6 * len(ext):].reshape(npts, 3)[:: max(1, npts // 50)],
            ref_x_cams=res.x[:6 * len(ext)], ref_wall_s=wall, versions=np.array([np.__version__, scipy.__version__]))
        print(name, costs, res.nfev, res.status, rms, lsmr_its, f"{wall:.1f} s", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["C2", "C4"])
