"""Device plan vs host plan on one full-size config (single GPU)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meatmodeler_b200 import _capi, synth
for name in sys.argv[1:] or ["C4"]:
    prob = synth.make_config(name, hard=True)
    ext, K, pts, uv, fi, pi = prob.args()
    nc, npts = len(ext), len(pts)
    host = _capi.plan(nc, npts, fi, pi)
    with _capi.Engine() as eng:
        for rep in range(2):
            eng.set_problem(nc, npts, K, fi, pi, uv)
            dev = eng.plan()
            bad = [k for k in ("n_tiles", "n_obs_local", "point_begin", "point_end") if dev[k] != host[k]] + [
                k for k in ("point_perm", "obs_perm", "meta", "tile_cams") if not np.array_equal(dev[k], host[k])]
            print(name, "rep", rep, "n_tiles", dev["n_tiles"], host["n_tiles"], "live", int((dev["obs_perm"] >= 0).sum()), "differs:", bad, flush=True)
            if "point_perm" in bad:
                d = np.flatnonzero(dev["point_perm"] != host["point_perm"])
                print("  first differing positions", d[:10], "dev", dev["point_perm"][d[:10]], "host", host["point_perm"][d[:10]],
                      "is permutation:", np.array_equal(np.sort(dev["point_perm"]), np.arange(npts)))
                fc = np.full(npts, nc); np.minimum.at(fc, pi, fi)
                print("  first cameras at those positions: dev", fc[dev["point_perm"][d[:10]]], "host", fc[host["point_perm"][d[:10]]])
