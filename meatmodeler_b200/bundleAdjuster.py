"""Drop-in replacement for MeatModeler's ``bundleAdjuster`` module, backed by ``libmmba.so``.

Put this directory in front of the reference on ``sys.path`` and ``processor.py`` runs unchanged:
``import bundleAdjuster`` (processor.py:8) then resolves here, and
``bundleAdjuster.adjustPoints(extrinsics, K, points_3D, points_2D, frame_indices, point_indices)``
(processor.py:465-470) keeps the reference signature and return value
(bundleAdjuster.py:160-194: ``(points (Np,3) float64, list of Nc 4x4 float64 extrinsics)``).

Only O(Nc) parameter packing/unpacking happens on the host (``frameParameters``,
``reformatPointResult`` — bundleAdjuster.py:105-157).  The residual model, the Jacobian and the
whole trust-region solve run in hand-written CUDA kernels behind the C-ABI of ``include/mmba.h``.
There is no CPU fallback: without ``libmmba.so`` or without an sm_100 GPU every solve raises.
"""
from __future__ import annotations

import atexit
import math
import os
import sys

import numpy as np

try:
    from meatmodeler_b200 import _capi
except ImportError:  # imported as a top-level module with only this directory on sys.path
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from meatmodeler_b200 import _capi

# Defaults of the reference call site (bundleAdjuster.py:180-192) and of scipy's least_squares.
FTOL = 1e-4
XTOL = 1e-8
GTOL = 1e-8
VERBOSE = 2          # the reference passes verbose=2: scipy prints its iteration table
TRACK_LIMIT = 256    # observations of one point (one tile of the streaming layout); longer tracks raise ValueError

#: statistics of the most recent solve (scipy ``OptimizeResult``-like fields)
last_result = None

_STATUS_MESSAGES = {
    0: "The maximum number of function evaluations is exceeded.",
    1: "`gtol` termination condition is satisfied.",
    2: "`ftol` termination condition is satisfied.",
    3: "`xtol` termination condition is satisfied.",
    4: "Both `ftol` and `xtol` termination conditions are satisfied.",
}


class SolveResult(dict):
    """Attribute-style result, the fields scipy's ``OptimizeResult`` carries for this call (missing attributes raise
    AttributeError, as OptimizeResult does, so that hasattr / getattr-with-default / copy / pickle behave)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__

    def __dir__(self):
        return list(self.keys())


class _SplitX:
    """Result parameters held as (cameras (Nc,6), points (Np,3)); ``np.asarray`` gives the packed vector."""

    def __init__(self, cams, points):
        self.cams, self.points = cams, points

    def __array__(self, dtype=None, copy=None):
        x = np.hstack((self.cams.reshape(-1), self.points.reshape(-1)))
        return x if dtype is None else x.astype(dtype)

    def __len__(self):
        return self.cams.size + self.points.size


def rotate(points, rot_vecs):
    """Rodrigues rotation of ``points[i]`` by ``rot_vecs[i]`` (bundleAdjuster.py:7-28), on the GPU."""
    return _capi.rotate(points, rot_vecs, device=_current_device())


def project(points, frame_params, camera_matrix):
    """Re-projection of ``points[i]`` with the parameters ``frame_params[i]`` (bundleAdjuster.py:31-52), on the GPU."""
    return _capi.project(points, frame_params, camera_matrix, device=_current_device())


def _current_device():
    if _is_distributed():
        import torch
        return torch.cuda.current_device()
    return 0


# --------------------------------------------------------------------------------------------------
# host-side packing (O(Nc) / O(Np) copies, bundleAdjuster.py:105-157)
# --------------------------------------------------------------------------------------------------

def frameParameters(frame_extrinsic_matrices):
    """Extrinsic matrices (Nc, 3 or 4, 4) -> flat ``[rvec0 | tvec0 | rvec1 | tvec1 ...]``.

    Same convention as bundleAdjuster.py:105-134: rotation angle from the trace, axis from the
    antisymmetric part divided by ``2 sin(angle)``, with 0/0 mapped to 0 (so the identity rotation
    gives a zero vector).
    """
    ext = np.asarray(frame_extrinsic_matrices, dtype=np.float64)
    R = ext[:, :3, :3]
    angle = np.arccos((np.trace(R, axis1=1, axis2=2) - 1.0) / 2.0)
    anti = np.stack((R[:, 2, 1] - R[:, 1, 2], R[:, 0, 2] - R[:, 2, 0], R[:, 1, 0] - R[:, 0, 1]), axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        axis = np.nan_to_num(anti / (2.0 * np.sin(angle))[:, None])
    out = np.empty((len(ext), 6))
    out[:, :3] = axis * angle[:, None]
    out[:, 3:] = ext[:, :3, 3]
    return out.reshape(-1)


def _rodrigues(rvecs):
    """Axis-angle vectors (N,3) -> rotations (N,3,3): the matrices ``cv2.Rodrigues`` returns
    (bundleAdjuster.py:153), all frames at once."""
    w = np.asarray(rvecs, dtype=np.float64).reshape(-1, 3)
    t = np.sqrt((w * w).sum(axis=1))
    safe = np.where(t > 0, t, 1.0)
    k = w / safe[:, None]
    Kx = np.zeros((len(w), 3, 3))
    Kx[:, 0, 1], Kx[:, 0, 2] = -k[:, 2], k[:, 1]
    Kx[:, 1, 0], Kx[:, 1, 2] = k[:, 2], -k[:, 0]
    Kx[:, 2, 0], Kx[:, 2, 1] = -k[:, 1], k[:, 0]
    c, s_ = np.cos(t)[:, None, None], np.sin(t)[:, None, None]
    R = c * np.eye(3)[None] + s_ * Kx + (1.0 - c) * (k[:, :, None] * k[:, None, :])
    R[t == 0] = np.eye(3)
    return R


def reformatPointResult(result, n_frames, n_points):
    """``result.x`` -> ((Np,3) points, list of Nc 4x4 extrinsics)   (bundleAdjuster.py:137-157)."""
    if isinstance(result.x, _SplitX):
        points, frames = result.x.points, result.x.cams
    else:
        x = np.asarray(result.x)
        points = x[n_frames * 6:].reshape((n_points, 3))
        frames = x[:n_frames * 6].reshape((n_frames, 6))
    ext = np.zeros((n_frames, 4, 4))
    ext[:, :3, :3] = _rodrigues(frames[:, :3])
    ext[:, :3, 3] = frames[:, 3:]
    ext[:, 3, 3] = 1.0
    return points, list(ext)


def pointAdjustmentSparsity(n_frames, n_points, frame_indices, point_indices):
    """Structural pattern of the Jacobian (bundleAdjuster.py:55-78) as a scipy CSR matrix.

    The engine never builds it (the block structure is implied by the two index arrays); it is kept
    for callers that want the pattern itself.
    """
    from scipy.sparse import csr_matrix

    fi = np.asarray(frame_indices, dtype=np.int64)
    pi = np.asarray(point_indices, dtype=np.int64)
    n_obs = fi.size
    cols = np.concatenate((6 * fi[:, None] + np.arange(6)[None, :],
                           6 * n_frames + 3 * pi[:, None] + np.arange(3)[None, :]), axis=1)
    cols = np.repeat(cols, 2, axis=0).reshape(-1)
    indptr = 9 * np.arange(2 * n_obs + 1)
    return csr_matrix((np.ones(cols.size, dtype=int), cols, indptr), shape=(2 * n_obs, 6 * n_frames + 3 * n_points))


# --------------------------------------------------------------------------------------------------
# engine access
# --------------------------------------------------------------------------------------------------

def _is_distributed():
    dist = getattr(sys.modules.get("torch"), "distributed", None)
    return bool(dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)


def _dist_options():
    """Shard options when the caller runs one process per GPU under ``torch.distributed``
    (observations sharded by point, cameras replicated, NCCL all-reduce inside the engine).
    A single process gets the single-GPU defaults; torch is not imported at all in that case."""
    if not _is_distributed():
        return {}
    import torch
    dist = torch.distributed
    rank, world = dist.get_rank(), dist.get_world_size()
    device = torch.cuda.current_device()
    blob = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{device}")
    if rank == 0:
        blob.copy_(torch.frombuffer(bytearray(_capi.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(blob, src=0)
    return dict(device=device, rank=rank, nranks=world, nccl_id=bytes(blob.cpu().numpy().tobytes()))


_ENGINES = {}          # (device, rank, nranks) -> live engine: one stream / NCCL communicator per process
_TUNABLE = ("ftol", "xtol", "gtol", "max_nfev", "pcg_rtol", "pcg_maxit", "verbose", "profile", "schur_mode", "pcg_atol", "pcg_ktol")


def _engine(camera_matrix, n_frames, n_points, frame_indices, point_indices, points_2D, **options):
    """The process-wide engine with this problem loaded.  Handles are cached: creating a CUDA stream,
    pinned staging buffers and (multi-GPU) an NCCL communicator once instead of once per call."""
    single = options.pop("single_gpu", False)
    fixed = {k: options.pop(k) for k in ("device", "rank", "nranks", "nccl_id") if k in options}
    unknown = [k for k in options if k not in _TUNABLE]
    if unknown:
        raise TypeError(f"unknown option(s) {unknown}")
    key = None
    if fixed:
        # explicitly configured engines are cached too (one per configuration), so that repeated calls do not leak a
        # device arena each; release() closes them
        key = ("explicit", fixed.get("device", 0), fixed.get("rank", 0), fixed.get("nranks", 1), bytes(fixed.get("nccl_id", b"")))
    if single and not fixed:
        # pose-only problems are tiny: every rank solves them whole on its own GPU
        dev = 0
        if _is_distributed():
            import torch
            dev = torch.cuda.current_device()
        key = (dev, 0, 1)
        fixed = dict(device=dev) if key not in _ENGINES else {}
    elif key is None:
        if _is_distributed():
            import torch
            key = (torch.cuda.current_device(), torch.distributed.get_rank(), torch.distributed.get_world_size())
        else:
            key = (0, 0, 1)
    eng = _ENGINES.get(key) if key is not None else None
    if eng is None:
        if not fixed and not single:
            fixed = _dist_options()
        eng = _capi.Engine(**fixed)
        if key is not None:
            _ENGINES[key] = eng
    defaults = _capi.default_options()
    eng.set_options(**{k: options.get(k, getattr(defaults, k)) for k in _TUNABLE})
    eng.set_problem(n_frames, n_points, camera_matrix, frame_indices, point_indices, points_2D)
    return eng


def release():
    """Destroy the cached engines (device memory, streams, communicators)."""
    for eng in _ENGINES.values():
        eng.close()
    _ENGINES.clear()


atexit.register(release)


def pointFun(parameters, camera_matrix, n_frames, n_points, frame_indices, point_indices, points_2D):
    """Reprojection residuals, interleaved (du0, dv0, du1, ...)  (bundleAdjuster.py:81-102),
    evaluated by the engine's residual kernel."""
    eng = _engine(camera_matrix, n_frames, n_points, frame_indices, point_indices, points_2D)
    return eng.residual(parameters)


def _print_table(log, res):
    """scipy's verbose=2 iteration table (scipy/optimize/_lsq/common.py:545-563) and summary."""
    print("{:^15}{:^15}{:^15}{:^15}{:^15}{:^15}".format(
        "Iteration", "Total nfev", "Cost", "Cost reduction", "Step norm", "Optimality"))
    for row in log:
        red = "" if math.isnan(row["cost_reduction"]) else f"{row['cost_reduction']:.2e}"
        stp = "" if math.isnan(row["step_norm"]) else f"{row['step_norm']:.2e}"
        print(f"{row['iteration']:^15}{row['nfev']:^15}{row['cost']:^15.4e}{red:^15}{stp:^15}"
              f"{row['optimality']:^15.2e}")
    print(_STATUS_MESSAGES.get(res.status, ""))
    print(f"Function evaluations {res.nfev}, initial cost {res.initial_cost:.4e}, final cost "
          f"{res.cost:.4e}, first-order optimality {res.optimality:.2e}.")


def solve(parameters, camera_matrix, n_frames, n_points, frame_indices, point_indices, points_2D,
          ftol=FTOL, xtol=XTOL, gtol=GTOL, max_nfev=None, verbose=0, want_fun=False, engine=None,
          **options):
    """The seam: stands where ``least_squares(pointFun, parameters, jac_sparsity=A, x_scale='jac',
    ftol=1e-4, method='trf', args=...)`` stands in the reference (bundleAdjuster.py:180-192).

    Returns a ``SolveResult`` with ``x, cost, fun, optimality, nfev, njev, nit, status, message,
    success`` plus engine statistics (``log``, ``pcg_iterations``, ``solve_ms``).
    """
    sharded = engine is None and _is_distributed()
    try:
        eng = engine if engine is not None else _engine(
            camera_matrix, n_frames, n_points, frame_indices, point_indices, points_2D,
            ftol=ftol, xtol=xtol, gtol=gtol, max_nfev=0 if max_nfev is None else int(max_nfev), **options)
    except _capi.MmbaError as e:
        if e.code == -5:
            # documented limit of the engine ("Known limits"): a point's observations live in one 256-slot
            # tile.  The reference has no such limit; the call fails loudly instead of solving a different problem.
            raise ValueError(f"adjustPoints: {e} - tracks of more than {TRACK_LIMIT} observations are not supported "
                             "by the B200 engine") from e
        raise
    try:
        if isinstance(parameters, tuple):
            # (camera parameters, points) as adjustPoints holds them: no packed host copy in either direction; the
            # packed vector scipy's OptimizeResult carries is available as ``res.x`` on demand (SolveResult.x)
            cams_out, points_out, r, fun = eng.solve_split(parameters[0], parameters[1], want_fun=want_fun)
            x = _SplitX(cams_out, points_out)
        else:
            x, r, fun = eng.solve(parameters, want_fun=want_fun)
    except _capi.MmbaError as e:
        if e.code == -4:   # scipy raises ValueError here (least_squares.py:945-946)
            raise ValueError("Residuals are not finite in the initial point.") from e
        raise
    log = eng.log()
    if sharded and fun is not None:
        # each rank holds the residuals of its own observations (zeros elsewhere)
        import torch
        t = torch.from_numpy(fun).cuda()
        torch.distributed.all_reduce(t)
        fun = t.cpu().numpy()
    res = SolveResult(x=x, cost=r.cost, initial_cost=r.initial_cost, fun=fun, optimality=r.optimality,
                      nfev=r.nfev, njev=r.njev, nit=r.nit, status=r.status,
                      message=_STATUS_MESSAGES.get(r.status, ""), success=r.status > 0, log=log,
                      pcg_iterations=r.pcg_iterations, solve_ms=r.solve_ms)
    if verbose >= 2:
        _print_table(log, res)
    elif verbose == 1:
        print(res.message)
    return res


def reformatPoseResult(result, n_frames):
    """``result.x`` (6 per frame) -> list of n_frames 3x4 extrinsics  (bundleAdjuster.py:197-203)."""
    frames = np.asarray(result.x)[:n_frames * 6].reshape((n_frames, 6))
    return list(np.concatenate((_rodrigues(frames[:, :3]), frames[:, 3:, None]), axis=2))


def _board_points(pattern_size):
    """The stationary chessboard of adjustPose (bundleAdjuster.py:220-223): a 4x3 grid (times 2) in
    the x-z plane, built in float32 like the reference (the values are small integers)."""
    points_3D = np.zeros((pattern_size, 3), np.float32)
    grid = np.mgrid[0:4, 0:3].T.reshape(-1, 2) * 2
    points_3D[:, 0] = grid[:, 0]
    points_3D[:, 2] = grid[:, 1]
    return points_3D


def _pose_problem(n_frames, frame_indices, point_indices, points_3D):
    """Engine view of a pose-only problem: every observation gets its own (constant) copy of its 3-D
    point, so that tiles stay point-aligned however many frames see one board corner."""
    fi = np.ascontiguousarray(frame_indices, dtype=np.int64).reshape(-1)
    pts = np.asarray(points_3D, dtype=np.float64)[np.asarray(point_indices, dtype=np.int64).reshape(-1)]
    return fi, np.arange(len(fi), dtype=np.int64), pts


def poseFun(parameters, camera_intrinsic_matrix, n_frames, frame_indices, point_indices, points_3D, points_2D):
    """Residuals of the pose-only problem (bundleAdjuster.py:206-211) from the engine's residual kernel."""
    fi, pi, pts = _pose_problem(n_frames, frame_indices, point_indices, points_3D)
    eng = _engine(camera_intrinsic_matrix, n_frames, len(pts), fi, pi, points_2D, single_gpu=True)
    return eng.residual(np.hstack((np.asarray(parameters, dtype=np.float64).reshape(-1), pts.reshape(-1))))


def solve_pose(parameters, camera_intrinsic_matrix, n_frames, frame_indices, point_indices, points_3D, points_2D,
               ftol=FTOL, xtol=XTOL, gtol=GTOL, max_nfev=None, verbose=0, want_fun=False):
    """Stands where ``least_squares(poseFun, parameters, verbose=2, ftol=1e-4, args=...)`` stands in the
    reference's adjustPose (bundleAdjuster.py:232-241): dense defaults (trf, exact trust-region step,
    x_scale=1), the 3-D points are constants."""
    fi, pi, pts = _pose_problem(n_frames, frame_indices, point_indices, points_3D)
    eng = _engine(camera_intrinsic_matrix, n_frames, len(pts), fi, pi, points_2D, single_gpu=True,
                  ftol=ftol, xtol=xtol, gtol=gtol, max_nfev=0 if max_nfev is None else int(max_nfev))
    x0 = np.hstack((np.asarray(parameters, dtype=np.float64).reshape(-1), pts.reshape(-1)))
    try:
        x, r, fun = eng.solve_pose(x0, want_fun=want_fun)
    except _capi.MmbaError as e:
        if e.code == -4:
            raise ValueError("Residuals are not finite in the initial point.") from e
        raise
    log = eng.log()
    res = SolveResult(x=x[:6 * n_frames], cost=r.cost, initial_cost=r.initial_cost, fun=fun, optimality=r.optimality,
                      nfev=r.nfev, njev=r.njev, nit=r.nit, status=r.status,
                      message=_STATUS_MESSAGES.get(r.status, ""), success=r.status > 0, log=log,
                      pcg_iterations=0, solve_ms=r.solve_ms)
    if verbose >= 2:
        _print_table(log, res)
    elif verbose == 1:
        print(res.message)
    return res


def adjustPose(frame_extrinsic_matrices, camera_intrinsic_matrix, points_2D):
    """Pose refinement against the fixed chessboard (bundleAdjuster.py:214-243; same signature and
    return value: a list of 3x4 extrinsic matrices).  ``points_2D`` holds n_frames copies of the
    board's image corners."""
    global last_result
    ext = np.asarray(frame_extrinsic_matrices, dtype=np.float64)
    n_frames = len(ext)
    pattern_size = int(len(points_2D) / n_frames)
    points_3D = _board_points(pattern_size)
    frame_indices = np.repeat(np.arange(n_frames), pattern_size)
    point_indices = np.repeat([np.arange(pattern_size)], n_frames, axis=0).reshape(pattern_size * n_frames)
    parameters = frameParameters(ext)
    res = solve_pose(parameters, camera_intrinsic_matrix, n_frames, frame_indices, point_indices, points_3D,
                     np.asarray(points_2D, dtype=np.float64).reshape(-1, 2), ftol=FTOL, verbose=VERBOSE)
    last_result = res
    return reformatPoseResult(res, n_frames)


def adjustPoints(frame_extrinsic_matrices, camera_intrinsic_matrix, points_3D, points_2D, frame_indices,
                 point_indices):
    """Bundle adjustment of all cameras and points (bundleAdjuster.py:160-194; same signature,
    same return value, same printed iteration table).

    :param frame_extrinsic_matrices: (Nc, 3|4, 4) extrinsic matrices
    :param camera_intrinsic_matrix: shared 3x3 intrinsic matrix
    :param points_3D: (Np, 3) or (Np, 1, 3) triangulated points
    :param points_2D: (No, 2) image observations
    :param frame_indices: (No,) camera of each observation
    :param point_indices: (No,) point of each observation
    :return: ((Np, 3) adjusted points, list of Nc 4x4 adjusted extrinsics)
    """
    global last_result
    ext = np.asarray(frame_extrinsic_matrices, dtype=np.float64)
    pts = np.asarray(points_3D, dtype=np.float64)
    n_frames, n_points = len(ext), len(pts)
    # the reference packs np.hstack((frameParameters(...), points)) (bundleAdjuster.py:172-176); the two halves go to
    # the engine as they are (mmba_solve_split), the packed vector is never materialised on the host
    res = solve((frameParameters(ext), pts.reshape((n_points * 3,))), camera_intrinsic_matrix, n_frames, n_points,
                frame_indices, point_indices, points_2D, ftol=FTOL, verbose=VERBOSE)
    last_result = res
    return reformatPointResult(res, n_frames, n_points)
