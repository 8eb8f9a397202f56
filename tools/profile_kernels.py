"""Launch each streaming kernel a few times on one config (for ncu captures and quick timings).

    python tools/profile_kernels.py [--config C2] [--iters 5] [--kernels matvec,jv,...]
Prints one line per kernel: average launch time (CUDA events on the engine's stream), algorithmic
GB/s and the fraction of the measured HBM peak.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import algorithmic_bytes, measured_peak  # noqa: E402
from meatmodeler_b200 import _capi, synth  # noqa: E402
from meatmodeler_b200 import bundleAdjuster as mm  # noqa: E402

CLASSES = {"build": _capi.K_BUILD, "resid": _capi.K_RESID, "schur_rhs": _capi.K_RHS, "schur_matvec": _capi.K_MATVEC,
           "backsub": _capi.K_BACKSUB, "jv": _capi.K_JV, "jv1": 100, "stream_read_J": 101,
           "schur_build": _capi.K_SBUILD, "schur_pcg": _capi.K_PCG}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--kernels", default="schur_matvec,jv,backsub,schur_rhs,build,resid")
    ap.add_argument("--random-visibility", action="store_true")
    args = ap.parse_args()
    prob = synth.make_config(args.config, hard=True, windowed=not args.random_visibility)
    ext, K, pts, uv, fi, pi = prob.args()
    nc, npts, nobs = prob.sizes
    x0 = np.hstack((mm.frameParameters(ext), pts.reshape(-1)))
    peak, kind = measured_peak()
    with _capi.Engine() as eng:
        eng.set_problem(nc, npts, K, fi, pi, uv)
        for name in args.kernels.split(","):
            ms = eng.bench_kernel(x0, CLASSES[name], args.iters)
            b = (algorithmic_bytes("jv", nc, npts, nobs) if name == "jv1" else 144 * 256 * eng.shard()["n_tiles"]
                 if name == "stream_read_J" else 0 if name == "schur_pcg" else algorithmic_bytes(name, nc, npts, nobs))
            print(json.dumps({"kernel": name, "config": args.config, "avg_ms": ms, "algorithmic_bytes": b,
                              "gbs": b / ms / 1e6, "frac_of_%s_peak" % kind: b / ms / 1e6 / peak}))


if __name__ == "__main__":
    main()
