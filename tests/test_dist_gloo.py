"""World-size-2 check of the sharding logic on the CPU (gloo): every rank takes its point range
from the tile plan (the same C++ code the engine uses), computes its shard's partial sums with the
oracle, all-reduces exactly the buffers the engine all-reduces with NCCL ([U | g_c | cost] after the
build pass, the 6*Nc Schur product per PCG iteration) and must reproduce the single-rank values."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from meatmodeler_b200 import _capi, synth
from oracle import ba_oracle as ba
from oracle import schur_trf


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prob = synth.make_problem(30, 700, 5000, seed=3, hard=True, windowed=False)
        ext, K, pts, uv, fi, pi = prob.args()
        nc, npts = len(ext), len(pts)
        x = np.hstack((ba.frame_parameters(ext), pts.reshape(-1)))
        pl = _capi.plan(nc, npts, fi, pi, rank, world)
        mine = pl["obs_perm"][pl["obs_perm"] >= 0]                     # this rank's observations
        own_pts = pl["point_perm"][pl["point_begin"]:pl["point_end"]]
        assert set(np.unique(pi[mine])) <= set(own_pts)                # a point's observations stay on one rank

        lin = schur_trf.Linearisation(x, K, nc, npts, fi[mine], pi[mine], uv[mine])
        buf = torch.from_numpy(np.hstack((lin.U.ravel(), lin.gc.ravel(), [2 * lin.cost])))
        dist.all_reduce(buf)                                           # [U | g_c | cost]
        U = buf[:36 * nc].numpy().reshape(nc, 6, 6)
        gc = buf[36 * nc:42 * nc].numpy().reshape(nc, 6)
        cost = 0.5 * float(buf[-1])

        # one implicit Schur product: y = sum_p W_p V'_p^-1 W_p^T v, shard-local then all-reduced
        reg = 1e-3
        v = np.random.default_rng(0).normal(size=(nc, 6))
        M = np.linalg.inv(lin.V + reg * np.eye(3)[None])               # unobserved points contribute nothing
        t = schur_trf._segsum(pi[mine], np.einsum("nij,ni->nj", lin.Jp, np.einsum("nij,nj->ni", lin.Jc, v[fi[mine]])), npts)
        z = np.einsum("nij,nj->ni", M, t)
        y = schur_trf._segsum(fi[mine], np.einsum("nij,ni->nj", lin.Jc, np.einsum("nij,nj->ni", lin.Jp, z[pi[mine]])), nc)
        yb = torch.from_numpy(y.copy())
        dist.all_reduce(yb)

        if rank == 0:
            full = schur_trf.Linearisation(x, K, nc, npts, fi, pi, uv)
            Mf = np.linalg.inv(full.V + reg * np.eye(3)[None])
            tf = schur_trf._segsum(pi, np.einsum("nij,ni->nj", full.Jp, np.einsum("nij,nj->ni", full.Jc, v[fi])), npts)
            zf = np.einsum("nij,nj->ni", Mf, tf)
            yf = schur_trf._segsum(fi, np.einsum("nij,ni->nj", full.Jc, np.einsum("nij,nj->ni", full.Jp, zf[pi])), nc)
            ok = (np.allclose(U, full.U, rtol=1e-12, atol=1e-9) and np.allclose(gc, full.gc, rtol=1e-12, atol=1e-9)
                  and abs(cost - full.cost) <= 1e-12 * full.cost and np.allclose(yb.numpy(), yf, rtol=1e-11, atol=1e-9))
            with open(os.path.join(out_dir, "result"), "w") as f:
                f.write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


def test_sharded_partial_sums_equal_whole(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "result").read_text() == "ok"


def test_unique_id_broadcast_plumbing_single_process():
    """Without torch.distributed initialised the drop-in must stay single-GPU (no NCCL, no torch import needed)."""
    from meatmodeler_b200 import bundleAdjuster as mm
    assert mm._dist_options() == {}
