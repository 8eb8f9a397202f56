"""Multi-GPU check, run under torchrun (one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
Every rank solves its point shard of the same problem (cameras replicated, NCCL all-reduce inside
libmmba); rank 0 additionally solves the whole problem on one GPU.  The sharded solve must agree
with the single-GPU solve: same iteration counts, per-iteration costs to 1e-7 (the bar is 1e-6), x to 1e-7.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from meatmodeler_b200 import _capi, synth  # noqa: E402
from meatmodeler_b200 import bundleAdjuster as mm  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    cases = [] if os.environ.get("MMBA_DIST_FULL") else None
    if cases is not None:
        cases = [("c4-full", synth.make_config("C4", hard=True))]
    for name, prob in cases or (("windowed", synth.make_problem(40, 3000, 24000, seed=21, hard=True)),
                       ("random", synth.make_problem(60, 1500, 9000, seed=7, hard=True, windowed=False)),
                       ("many-cameras", synth.make_problem(300, 900, 4000, seed=9, windowed=False)),   # cluster update path
                       ("c2-tenth", synth.make_config("C2", hard=True, scale=0.1)),
                       ("c4-tenth", synth.make_config("C4", hard=True, scale=0.1))):
        ext, K, pts, uv, fi, pi = prob.args()
        nc, npts = len(ext), len(pts)
        x0 = np.hstack((mm.frameParameters(ext), np.asarray(pts).reshape(-1)))
        res = mm.solve(x0, K, nc, npts, fi, pi, uv, want_fun=True)       # sharded (torch.distributed is initialised)
        costs = np.array([r["cost"] for r in res.log])
        # every rank: the shard plan built on the device (chunked upload + all-to-all) == the host plan of this rank
        eng_s = next(iter(mm._ENGINES.values()))
        dev, host = eng_s.plan(), _capi.plan(nc, npts, fi, pi, rank, world)
        cnt, fst, lst, fhi = eng_s.plan_stats()
        cnt_ref = np.bincount(pi, minlength=npts)
        fst_ref = np.full(npts, nc); np.minimum.at(fst_ref, pi, fi)
        lst_ref = np.zeros(npts, dtype=np.int64); np.maximum.at(lst_ref, pi, fi)     # (the engine starts the max at 0)
        wrong = np.flatnonzero((cnt != cnt_ref) | (fst != fst_ref) | (lst != lst_ref))
        if len(wrong):
            w = wrong[:6]
            first_obs = np.searchsorted(pi, w)
            print(f"{name}: rank {rank} stats wrong for {len(wrong)} points, e.g. {w} count {cnt[w]} ref {cnt_ref[w]} first {fst[w]} ref {fst_ref[w]} "
                  f"last {lst[w]} ref {lst_ref[w]}; first obs index of those points {first_obs}; range of wrong points {wrong.min()}..{wrong.max()}",
                  flush=True)
        bad = [k for k in ("n_tiles", "n_obs_local", "point_begin", "point_end") if dev[k] != host[k]] + [
            k for k in ("point_perm", "obs_perm", "meta", "tile_cams") if not np.array_equal(dev[k], host[k])]
        plan_ok = not bad
        print(f"{name}: rank {rank} n_obs_local {dev['n_obs_local']} n_tiles {dev['n_tiles']} live slots {int((dev['obs_perm'] >= 0).sum())}", flush=True)
        if bad:
            print(f"{name}: rank {rank} plan differs in {bad}: device {[dev[k] for k in bad if np.isscalar(dev[k])]} "
                  f"host {[host[k] for k in bad if np.isscalar(host[k])]}", flush=True)
        t_ok = torch.tensor([1 if plan_ok else 0], device="cuda")
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"{name}: world={world} shard plans device == host -> {'OK' if t_ok.item() else 'MISMATCH'}", flush=True)
            ok = ok and bool(t_ok.item())
        if rank == 0:
            with _capi.Engine(device=local) as eng:                        # single GPU
                eng.set_problem(nc, npts, K, fi, pi, uv)
                x1, r1, f1 = eng.solve(x0, want_fun=True)
                c1 = np.array([r["cost"] for r in eng.log()])
            # (x itself is loosely determined along the gauge directions of the long camera chains: 1e-5 there)
            same = (res.nfev == r1.nfev and res.status == r1.status and len(costs) == len(c1)
                    and np.allclose(costs, c1, rtol=1e-7) and np.abs(res.x - x1).max() < (1e-5 if nc > 1000 else 1e-7)
                    and np.abs(res.fun - f1).max() < 1e-5)
            if not same:
                df = np.abs(res.fun - f1).reshape(-1, 2).max(axis=1)
                badobs = np.flatnonzero(df > 1e-5)
                print(f"{name}: costs sharded {costs} single {c1}; {len(badobs)} observations with |df| > 1e-5, first {badobs[:8]}, "
                      f"zero residuals in sharded fun: {int((np.abs(res.fun).reshape(-1, 2).max(axis=1) == 0).sum())}", flush=True)
            print(f"{name}: world={world} nfev {res.nfev}/{r1.nfev} status {res.status}/{r1.status} "
                  f"cost {res.cost:.12e}/{r1.cost:.12e} max|dx| {np.abs(res.x - x1).max():.2e} "
                  f"max|df| {np.abs(res.fun - f1).max():.2e} pcg {res.pcg_iterations}/{r1.pcg_iterations} -> {'OK' if same else 'MISMATCH'}",
                  flush=True)
            ok = ok and same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
