#!/usr/bin/env python
"""Benchmark of the bundle-adjustment hot path (BASELINE.json metric: BA observations/s, one LM
iteration = residual + Jacobian + Schur/PCG solve + step selection).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl mmba|reference] [--config auto|C1..C5]

A "step" is one complete solve (the work of one ``adjustPoints`` call) from the same synthetic
starting point; ``value`` = observations x LM iterations of the K timed steps / device time, with
the problem resident in HBM; ``e2e`` = the same metric through the drop-in
``bundleAdjuster.adjustPoints`` call with host (numpy) buffers, i.e. including the host-side plan,
every host->device copy and the device->host read of the result.

Workloads: N=1 -> BASELINE configs[1] (200 cameras, 50k points, 1M observations).  N>1 -> the same
shape per GPU (200 cameras replicated, N x 50k points, N x 1M observations sharded by point):
"scaling": "weak".  ``--config C4`` runs the 5M-observation Venice-sized problem at any N (strong).

``--impl reference`` times the reference's CPU path (scipy least_squares with the reference's
arguments, restated in oracle/ba_oracle.py because the reference is a Python module that cannot be
compiled into oracle/_ref) on a bounded sample of the same workload, on rank 0 only.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "BA observations/sec (resid+Jacobian+Schur per LM iter)"
UNIT = "observations/s"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def workload(name, n_gpus):
    from meatmodeler_b200 import synth
    if name == "auto":
        nc, npts, nobs = synth.CONFIGS["C2"]
        prob = synth.make_problem(nc, npts * n_gpus, nobs * n_gpus, seed=synth.CONFIG_SEEDS["C2"], hard=True)
        label = "C2 food-video shape: 200 cameras, 50k points, 1M observations" + (
            f" per GPU x {n_gpus} (points and observations scaled, cameras replicated)" if n_gpus > 1 else "")
        return prob, label, "weak"
    prob = synth.make_config(name, hard=True)
    nc, npts, nobs = prob.sizes
    return prob, f"{name}: {nc} cameras, {npts} points, {nobs} observations", "strong"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons of one GPU, sampled every 200 ms while running."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_sample(scale=None):
    """The reference's CPU path on a bounded sample: C2 scaled down (all 200 cameras kept, points
    and observations x scale), full solve with the reference's least_squares arguments."""
    from threadpoolctl import threadpool_info
    from meatmodeler_b200 import synth
    from oracle import ba_oracle as ba        # CPU baseline leg: the one place bench.py runs oracle/

    scale = 0.03 if scale is None else scale
    prob = synth.make_config("C2", hard=True, scale=scale)
    ext, K, pts, uv, fi, pi = prob.args()
    rec = []
    t0 = time.perf_counter()
    res = ba.solve_reference_path(ext, K, pts, uv, fi, pi, record=rec)
    wall = time.perf_counter() - t0
    nit = len(rec)
    threads = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    return {"value": len(uv) * nit / wall, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"C2 scaled x{scale}: {prob.sizes[0]} cameras, {prob.sizes[1]} points, {prob.sizes[2]} observations, "
                      f"{nit} LM iterations (nfev {res.nfev}) in {wall:.2f} s; scipy TRF+LSMR with 2-point sparse finite "
                      f"differences, numpy/scipy effectively single-threaded (BLAS threads available: {threads}, "
                      f"host cpus: {os.cpu_count()})",
            "wall_s": wall, "lm_iterations": nit, "cost": float(res.cost), "n_obs": len(uv)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prob, label, scaling = workload(args.config, args.gpus) if args.config != "auto" else (None, None, "weak")
    if label is None:
        label = "C2 food-video shape: 200 cameras, 50k points, 1M observations" + (
            f" per GPU x {args.gpus} (points and observations scaled, cameras replicated)" if args.gpus > 1 else "")
    del prob
    samples = []
    for i in range(args.warmup + args.steps):
        s = cpu_reference_sample(scale=args.ref_scale)
        if i >= args.warmup:
            samples.append(s)
    wall = sum(s["wall_s"] for s in samples)
    its = sum(s["lm_iterations"] for s in samples)
    nobs = samples[0]["n_obs"]
    value = nobs * its / wall
    base = dict(samples[-1])
    base["value"] = value
    for k in ("wall_s", "lm_iterations", "cost", "n_obs"):
        base.pop(k, None)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / len(samples), "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": label, "timed_on": "bounded CPU sample of the workload, see cpu_baseline.sample"},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def algorithmic_bytes(kernel, nc, npts, nobs, nnz_up=0):
    """SURVEY.md §8d / DESIGN.md: bytes one launch must move (6-parameter cameras)."""
    return {
        # explicit reduced camera matrix: J + tile metadata in, M and M g_p per point in, y and the upper blocks out
        "schur_build": 152 * nobs + 72 * npts + 48 * nc + 288 * nnz_up,
        "build": 184 * nobs + 96 * npts + 264 * nc,
        "schur_matvec": 152 * nobs + 48 * npts + 96 * nc,
        "schur_rhs": 152 * nobs + 72 * npts + 264 * nc,
        "backsub": 152 * nobs + 96 * npts + 48 * nc,
        "jv": 152 * nobs + 2 * (24 * npts + 48 * nc),
        "resid": 24 * nobs + 24 * npts + 48 * nc,
    }[kernel]


def run_mmba(args):
    import torch
    import torch.distributed as dist

    from meatmodeler_b200 import _capi
    from meatmodeler_b200 import bundleAdjuster as mm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    prob, label, scaling = workload(args.config, args.gpus)
    ext, K, pts, uv, fi, pi = prob.args()
    nc, npts, nobs = prob.sizes
    x0 = np.hstack((mm.frameParameters(ext), pts.reshape(-1)))

    schur_mode = {"auto": _capi.SCHUR_AUTO, "implicit": _capi.SCHUR_IMPLICIT, "explicit": _capi.SCHUR_EXPLICIT}[args.schur]
    opts = mm._dist_options()
    opts.setdefault("device", local)
    opts["schur_mode"] = schur_mode
    eng = _capi.Engine(**opts)
    t0 = time.perf_counter()
    eng.set_problem(nc, npts, K, fi, pi, uv)
    setup_ms = 1e3 * (time.perf_counter() - t0)
    shard = eng.shard()
    eng.set_x(x0)

    # ---- device-resident timing: W warm-up + K timed solves -----------------------------------
    for _ in range(args.warmup):
        eng.solve_resident()
    barrier()
    dev_ms, its, pcg, launches = 0.0, 0, 0, 0
    with ClockSampler(local) as clocks:
        wall0 = time.perf_counter()
        for _ in range(args.steps):
            r = eng.solve_resident()
            dev_ms += r.solve_ms
            its += r.nit
            pcg += r.pcg_iterations
            launches += sum(v["launches"] for k, v in eng.profile().items() if k != "allreduce")
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - wall0)
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)
    value = nobs * its / (dev_ms * 1e-3)
    log = eng.log()
    final_cost, nit_per_step = r.cost, r.nit

    # ---- per-kernel durations: one extra step with every launch bracketed by CUDA events --------
    eng.close()
    opts_p = mm._dist_options()          # a fresh ncclUniqueId: one id initialises one communicator
    opts_p.setdefault("device", local)
    opts_p["profile"] = 1
    opts_p["schur_mode"] = schur_mode
    engp = _capi.Engine(**opts_p)
    engp.set_problem(nc, npts, K, fi, pi, uv)
    engp.set_x(x0)
    engp.solve_resident()
    rp = engp.solve_resident()
    prof = engp.profile()
    engp.close()
    peak, peak_kind = measured_peak()
    n_obs_local, n_pts_local = shard["n_obs_local"], shard["n_points_local"]
    kernels = {}
    explicit = prof["schur_pcg"]["launches"] > 0
    nnz_up = nnz_full = total_pairs = 0
    if explicit:
        _, up_cols, nnz_full, total_pairs = _capi.host_rcm_pattern(nc, npts, fi, pi)
        nnz_up = len(up_cols)
    for name in ("build", "schur_build", "schur_matvec", "schur_rhs", "backsub", "jv", "resid"):
        p = prof[name]
        if p["launches"]:
            avg_ms = p["ms"] / p["launches"]
            b = algorithmic_bytes(name, nc, n_pts_local, n_obs_local, nnz_up)
            kernels[name] = {"launches_per_step": p["launches"], "avg_ms": avg_ms, "share_of_step": p["ms"] / rp.solve_ms,
                             "algorithmic_bytes": b, "gbs": b / (avg_ms * 1e-3) / 1e9, "frac": b / (avg_ms * 1e-3) / 1e9 / peak}
    dom = max(kernels, key=lambda k: kernels[k]["share_of_step"])
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(f"{dom}:{args.config}:{args.gpus}")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "traffic": traffic, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
                "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"], "avg_launch_ms": kernels[dom]["avg_ms"],
                "how": "CUDA events around every launch on the engine's stream during one extra, identical step "
                       "(profile=1); the timed steps carry no per-launch events",
                "kernels": kernels,
                "other_classes_ms_per_step": {k: {"launches": prof[k]["launches"], "ms": prof[k]["ms"]}
                                              for k in ("vec", "allreduce", "cam_prep", "point_invert")},
                "profiled_step_ms": rp.solve_ms}
    if explicit and "schur_build" in kernels:
        # The S-build pass streams J once (HBM figure above) but is bound by FP64 issue: 108 DFMA per observation pair
        # of a point (sum over points of L (L + 1) / 2 pairs; this rank's share for a sharded solve).
        kb = kernels["schur_build"]
        flops = 2.0 * 108.0 * total_pairs * (n_obs_local / max(nobs, 1))
        kb["fp64"] = {"algorithmic_dfma": flops / 2, "achieved_tflops": flops / (kb["avg_ms"] * 1e-3) / 1e12,
                      "nominal_peak_tflops": 37.2, "frac_of_nominal": flops / (kb["avg_ms"] * 1e-3) / 1e12 / 37.2,
                      "peak_source": "148 SMs x 64 DFMA/clk x 2 x 1.965 GHz (vector FP64, nominal)"}
        if dom == "schur_build":
            roofline["note"] = ("the dominant streaming kernel (S-build) is FP64-issue bound, not HBM bound: see "
                                "kernels.schur_build.fp64; the HBM-bound passes are build / backsub / jv / resid")
    if explicit:
        # The PCG on the explicit, L2/L1-resident reduced camera matrix is one cooperative launch per outer
        # iteration: no HBM stream, two grid-wide exchanges (self-validating lines through L2) per PCG iteration.  It
        # has no HBM roofline; what bounds it is the L2 round-trip latency, reported as time per PCG iteration.
        p = prof["schur_pcg"]
        its_p = max(int(rp.pcg_iterations), 1)
        roofline["schur_pcg_on_chip"] = {
            "launches_per_step": p["launches"], "ms_per_step": p["ms"], "share_of_step": p["ms"] / rp.solve_ms,
            "pcg_iterations_per_step": int(rp.pcg_iterations), "us_per_pcg_iteration": 1e3 * p["ms"] / its_p,
            "matrix_bytes": 288 * nnz_full, "blocks_full": nnz_full, "blocks_upper": nnz_up,
            "l2_gbs": 288 * nnz_full * its_p / (p["ms"] * 1e-3) / 1e9 if p["ms"] > 0 else None,
            "bound": "latency of two grid-wide exchanges through L2 per iteration, not HBM"}

    # ---- end to end through the drop-in adjustPoints with host buffers ---------------------------
    e2e_steps = max(2, min(args.steps, 5))
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        mm.adjustPoints(ext, K, pts, uv, fi, pi)        # warm-up
    barrier()
    t0 = time.perf_counter()
    e2e_its = 0
    for _ in range(e2e_steps):
        with contextlib.redirect_stdout(sink):
            out_pts, out_ext = mm.adjustPoints(ext, K, pts, uv, fi, pi)
        e2e_its += mm.last_result.nit
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    ns = shard["n_tiles"] * 256
    h2d = 16 * ns + 8 * ns + 16 * shard["n_tiles"] + 8 * (6 * nc + 3 * n_pts_local)
    d2h = 8 * (6 * nc + 3 * n_pts_local)
    e2e = {"value": nobs * e2e_its / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps,
           "includes": "frameParameters, tile plan (host), all H2D copies, solve, D2H of x, reformatPointResult"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": label, "cameras": nc, "points": npts, "observations": nobs,
                       "init": "hard (points sigma 0.15, tvec sigma 0.1), 0.5 px noise, windowed visibility",
                       "l2": f"working set {(18 + 2 + 2 + 1) * 8 * n_obs_local / 1e6:.0f} MB per GPU streamed per pass "
                             f"(> 126 MB L2)" if n_obs_local * 184 > 126e6 else "working set fits L2 (no flush)",
                       "ftol": 1e-4, "pcg_rtol": eng.options.pcg_rtol,
                       "schur": ("explicit reduced camera matrix + one-kernel PCG" if explicit else
                                 "implicit (one streaming pass over J per PCG iteration)")},
            "lm_iterations_per_step": nit_per_step, "lm_iterations_per_s": its / (dev_ms * 1e-3),
            "pcg_iterations_per_step": pcg / args.steps, "final_cost": final_cost,
            "wall_ms_per_step": wall_ms / args.steps, "setup_ms": setup_ms,
            "gpu_launches": launches, "clocks": clocks.summary(), "e2e": e2e, "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = {k: v for k, v in cpu_reference_sample(scale=args.ref_scale).items()
                                    if k not in ("wall_s", "lm_iterations", "cost", "n_obs")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mmba", choices=["mmba", "reference"])
    ap.add_argument("--config", default="auto", choices=["auto", "C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--ref-scale", type=float, default=None, help="size of the CPU sample relative to C2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--schur", default="auto", choices=["auto", "implicit", "explicit"],
                    help="reduced camera system: explicit block-sparse matrix or implicit products (auto: library rule)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_mmba(args)


if __name__ == "__main__":
    main()
