"""ctypes binding of ``libmmba.so`` (C-ABI: ``include/mmba.h``).

This is the whole host<->engine boundary: plain pointers and sizes, caller-owned numpy buffers.
There is no CPU fallback — if the library is missing or no sm_100 device is usable the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

# MMBA_LIB: another build of the same library (e.g. the -DMMBA_PHASE_TIMING diagnostics build)
_LIB_PATH = os.environ.get("MMBA_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmmba.so")

K_NAMES = ("cam_prep", "build", "resid", "point_invert", "schur_rhs", "schur_matvec", "backsub", "jv",
           "vec", "allreduce", "schur_build", "schur_pcg")
(K_CAMPREP, K_BUILD, K_RESID, K_PTINV, K_RHS, K_MATVEC, K_BACKSUB, K_JV, K_VEC, K_ALLREDUCE, K_SBUILD,
 K_PCG) = range(12)
K_COUNT = 12
SCHUR_AUTO, SCHUR_IMPLICIT, SCHUR_EXPLICIT = 0, 1, 2
TILE_META_BYTES = 2592      # sizeof(TileMeta), csrc/plan.h

ERR_NAMES = {-1: "MMBA_ERR_ARG", -2: "MMBA_ERR_CUDA", -3: "MMBA_ERR_STATE", -4: "MMBA_ERR_NONFINITE",
             -5: "MMBA_ERR_TRACK", -6: "MMBA_ERR_NCCL", -7: "MMBA_ERR_NOMEM"}


class MmbaError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {message}")
        self.code = code


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32), ("verbose", C.c_int32),
                ("ftol", C.c_double), ("xtol", C.c_double), ("gtol", C.c_double), ("max_nfev", C.c_int64),
                ("pcg_rtol", C.c_double), ("pcg_maxit", C.c_int32), ("profile", C.c_int32),
                ("nccl_id", C.c_uint8 * 128), ("schur_mode", C.c_int32), ("reserved", C.c_int32),
                ("pcg_atol", C.c_double), ("pcg_ktol", C.c_double)]


class Result(C.Structure):
    _fields_ = [("cost", C.c_double), ("initial_cost", C.c_double), ("optimality", C.c_double),
                ("nfev", C.c_int64), ("njev", C.c_int64), ("nit", C.c_int64), ("status", C.c_int32),
                ("reserved", C.c_int32), ("pcg_iterations", C.c_int64), ("solve_ms", C.c_double)]


class IterLog(C.Structure):
    _fields_ = [("iteration", C.c_int64), ("nfev", C.c_int64), ("cost", C.c_double),
                ("cost_reduction", C.c_double), ("step_norm", C.c_double), ("optimality", C.c_double),
                ("reg", C.c_double), ("delta", C.c_double), ("pcg_iterations", C.c_int64)]


_f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64 = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_H = C.c_void_p

# every symbol include/mmba.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "mmba_version": (C.c_int, []),
    "mmba_last_error": (C.c_char_p, [_H]),
    "mmba_default_options": (None, [C.POINTER(Options)]),
    "mmba_create": (C.c_int, [C.POINTER(_H), C.POINTER(Options)]),
    "mmba_nccl_unique_id": (C.c_int, [C.POINTER(C.c_uint8 * 128)]),
    "mmba_destroy": (None, [_H]),
    "mmba_set_options": (C.c_int, [_H, C.POINTER(Options)]),
    "mmba_set_problem": (C.c_int, [_H, C.c_int64, C.c_int64, C.c_int64, _f64, _i64, _i64, _f64]),
    "mmba_solve": (C.c_int, [_H, _f64, C.POINTER(Result), C.c_void_p]),
    "mmba_solve_split": (C.c_int, [_H, _f64, _f64, _f64, _f64, C.POINTER(Result), C.c_void_p]),
    "mmba_solve_pose": (C.c_int, [_H, _f64, C.POINTER(Result), C.c_void_p]),
    "mmba_set_x": (C.c_int, [_H, _f64]),
    "mmba_solve_resident": (C.c_int, [_H, C.POINTER(Result)]),
    "mmba_get_x": (C.c_int, [_H, _f64]),
    "mmba_get_log": (C.c_int, [_H, C.POINTER(IterLog), C.c_int]),
    "mmba_get_profile": (C.c_int, [_H, C.POINTER(C.c_int64 * K_COUNT), C.POINTER(C.c_double * K_COUNT)]),
    "mmba_get_pcg_history": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_int]),
    "mmba_get_phase_cycles": (C.c_int, [_H, C.POINTER(C.c_int64 * 64), C.c_int]),
    "mmba_get_shard": (C.c_int, [_H, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mmba_eval_residual": (C.c_int, [_H, _f64, _f64]),
    "mmba_eval_jacobian": (C.c_int, [_H, _f64, _f64, _f64]),
    "mmba_eval_blocks": (C.c_int, [_H, _f64, _f64, _f64, _f64, _f64, C.POINTER(C.c_double)]),
    "mmba_eval_gn_step": (C.c_int, [_H, _f64, _f64, C.c_double, _f64, C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "mmba_eval_reduced_system": (C.c_int, [_H, _f64, _f64, C.c_double, _f64, _f64]),
    "mmba_eval_jnorm2": (C.c_int, [_H, _f64, _f64, C.POINTER(C.c_double)]),
    "mmba_bench_kernel": (C.c_int, [_H, _f64, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "mmba_triangulate": (C.c_int, [C.c_int, C.c_int64, _f64, C.c_int64, _i64, _i64, _f64, _f64, _f64, C.POINTER(C.c_double)]),
    "mmba_rotate": (C.c_int, [C.c_int, C.c_int64, _f64, _f64, _f64]),
    "mmba_project": (C.c_int, [C.c_int, C.c_int64, _f64, _f64, C.c_int64, _f64, _f64]),
    "mmba_host_tr2d": (C.c_int, [C.POINTER(C.c_double * 3), C.POINTER(C.c_double * 2), C.c_double,
                                 C.POINTER(C.c_double * 2), C.POINTER(C.c_int)]),
    "mmba_host_min_quadratic_1d": (C.c_int, [C.c_double, C.c_double, C.c_double, C.c_double,
                                             C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mmba_host_update_tr_radius": (C.c_int, [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                             C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mmba_host_check_termination": (C.c_int, [C.c_double] * 7),
    "mmba_host_rcm_pattern": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, _i64, _i64, C.POINTER(C.c_int64 * 3), C.c_void_p,
                                        C.c_void_p, C.c_int64]),
    "mmba_plan_create": (C.c_int, [C.POINTER(_H), C.c_int64, C.c_int64, C.c_int64, _i64, _i64, C.c_int, C.c_int]),
    "mmba_plan_destroy": (None, [_H]),
    "mmba_plan_sizes": (C.c_int, [_H, C.POINTER(C.c_int64 * 8)]),
    "mmba_plan_export": (C.c_int, [_H, _i64, _i64, _i32, _i32]),
    "mmba_plan_tile_stats": (C.c_int, [_H, _i32, _i32, _i32, _i32]),
    "mmba_plan_raw": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "mmba_get_plan_sizes": (C.c_int, [_H, C.POINTER(C.c_int64 * 8)]),
    "mmba_get_plan_raw": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mmba_get_plan_stats": (C.c_int, [_H, _i32, _i32, _i32, _i32]),
    "mmba_get_rcm_pattern": (C.c_int, [_H, C.POINTER(C.c_int64 * 8), C.c_void_p, C.c_void_p, C.c_int64]),
}

_lib = None


def lib():
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                              "g.build()'` (or `make -C meatmodeler_b200/csrc`); there is no CPU fallback")
        handle = C.CDLL(_LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def _check(code, handle=None):
    if code < 0:
        msg = lib().mmba_last_error(handle)
        raise MmbaError(code, msg.decode() if msg else "")
    return code


def default_options() -> Options:
    opt = Options()
    lib().mmba_default_options(C.byref(opt))
    return opt


def nccl_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    _check(lib().mmba_nccl_unique_id(C.byref(buf)))
    return bytes(buf)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Engine:
    """One ``mmba_handle``: a problem resident on one GPU (or one shard of it on one rank)."""

    def __init__(self, options: Options | None = None, **kw):
        opt = options if options is not None else default_options()
        nccl_id = kw.pop("nccl_id", None)
        for k, v in kw.items():
            if not hasattr(opt, k):
                raise TypeError(f"unknown option {k}")
            setattr(opt, k, v)
        if nccl_id is not None:
            C.memmove(opt.nccl_id, bytes(nccl_id), 128)
        self.options = opt
        self._h = _H()
        _check(lib().mmba_create(C.byref(self._h), C.byref(opt)))
        self.sizes = None

    def set_options(self, **kw):
        """Change tolerances / limits of the live handle (ftol, xtol, gtol, max_nfev, pcg_rtol, pcg_maxit,
        verbose, profile, schur_mode — the latter is read by the next set_problem; IMPLICIT applies at once)."""
        for k, v in kw.items():
            if k in ("device", "rank", "nranks", "nccl_id") or not hasattr(self.options, k):
                raise TypeError(f"option {k} cannot be changed on a live engine")
            setattr(self.options, k, v)
        _check(lib().mmba_set_options(self._h, C.byref(self.options)), self._h)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().mmba_destroy(self._h)
            self._h = _H()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- problem ------------------------------------------------------------------------------
    def set_problem(self, n_cams, n_points, K, cam_idx, pt_idx, uv):
        cam_idx = _c(cam_idx, np.int64).reshape(-1)
        pt_idx = _c(pt_idx, np.int64).reshape(-1)
        uv = _c(uv, np.float64).reshape(-1, 2)
        if not (len(cam_idx) == len(pt_idx) == len(uv)):
            raise ValueError("cam_idx, pt_idx and uv must have one entry per observation")
        K = _c(K, np.float64).reshape(9)
        _check(lib().mmba_set_problem(self._h, int(n_cams), int(n_points), len(uv), K, cam_idx, pt_idx, uv), self._h)
        self.sizes = (int(n_cams), int(n_points), len(uv))

    @property
    def n(self):
        return 6 * self.sizes[0] + 3 * self.sizes[1]

    def shard(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _check(lib().mmba_get_shard(self._h, C.byref(a), C.byref(b), C.byref(c)), self._h)
        return dict(n_obs_local=a.value, n_points_local=b.value, n_tiles=c.value)

    def plan(self):
        """The tile plan built on the device, in the format of ``_capi.plan`` (host builder) plus the raw records."""
        sizes = (C.c_int64 * 8)()
        _check(lib().mmba_get_plan_sizes(self._h, C.byref(sizes)), self._h)
        keys = ("n_tiles", "n_obs_local", "n_points_local", "point_begin", "point_end", "tile_obs", "max_tile_cams", "n_slots")
        out = dict(zip(keys, list(sizes)))
        nt, ns = out["n_tiles"], out["n_slots"]
        meta = np.zeros((max(nt, 1), TILE_META_BYTES), dtype=np.uint8)
        tile_cams = np.full((max(nt, 1), 256), -1, dtype=np.int32)
        obs_perm = np.full(max(ns, 1), -1, dtype=np.int64)
        point_perm = np.zeros(self.sizes[1], dtype=np.int64)
        _check(lib().mmba_get_plan_raw(self._h, meta.ctypes.data, tile_cams.ctypes.data, obs_perm.ctypes.data,
                                       point_perm.ctypes.data), self._h)
        out.update(meta=meta[:nt], tile_cams=tile_cams[:nt], obs_perm=obs_perm[:ns], point_perm=point_perm)
        return out

    def plan_stats(self):
        """Per-point statistics of the device plan: (count, first camera, last camera, first camera of the upper half)."""
        out = [np.zeros(self.sizes[1], dtype=np.int32) for _ in range(4)]
        _check(lib().mmba_get_plan_stats(self._h, *out), self._h)
        return out

    def rcm_pattern(self):
        """Device-built block pattern: dict(up_rowptr, up_cols, nnz_full, total_pairs, n_ctas, cpc, nblk_max, nh_max)."""
        sizes = (C.c_int64 * 8)()
        _check(lib().mmba_get_rcm_pattern(self._h, C.byref(sizes), None, None, 0), self._h)
        rowptr = np.empty(self.sizes[0] + 1, dtype=np.int32)
        cols = np.empty(max(int(sizes[0]), 1), dtype=np.int32)
        _check(lib().mmba_get_rcm_pattern(self._h, C.byref(sizes), rowptr.ctypes.data, cols.ctypes.data, len(cols)), self._h)
        return dict(up_rowptr=rowptr, up_cols=cols[:sizes[0]], nnz_full=int(sizes[1]), total_pairs=int(sizes[2]),
                    n_ctas=int(sizes[3]), cpc=int(sizes[4]), nblk_max=int(sizes[5]), nh_max=int(sizes[6]))

    # -- solve --------------------------------------------------------------------------------
    def solve(self, x0, want_fun=False):
        """Returns (x, Result, fun or None).  x0 is not modified."""
        x = np.array(x0, dtype=np.float64, order="C").reshape(-1)
        if x.size != self.n:
            raise ValueError(f"x0 has {x.size} entries, expected {self.n}")
        res = Result()
        fun = np.empty(2 * self.sizes[2]) if want_fun else None
        _check(lib().mmba_solve(self._h, x, C.byref(res), fun.ctypes.data if want_fun else None), self._h)
        return x, res, fun

    def solve_split(self, cams, points, want_fun=False):
        """The solve with cameras (6 per camera) and points (3 per point) in separate arrays, as adjustPoints holds them:
        returns (cams_out (Nc,6), points_out (Np,3), Result, fun or None); inputs are read in place, never copied or modified."""
        cams = _c(cams, np.float64).reshape(-1)
        points = _c(points, np.float64).reshape(-1)
        if cams.size != 6 * self.sizes[0] or points.size != 3 * self.sizes[1]:
            raise ValueError("cams / points do not match the problem's sizes")
        cams_out = np.empty((self.sizes[0], 6))
        points_out = np.empty((self.sizes[1], 3))
        res = Result()
        fun = np.empty(2 * self.sizes[2]) if want_fun else None
        _check(lib().mmba_solve_split(self._h, cams, points, cams_out.reshape(-1), points_out.reshape(-1), C.byref(res),
                                      fun.ctypes.data if want_fun else None), self._h)
        return cams_out, points_out, res, fun

    def solve_pose(self, x0, want_fun=False):
        """Pose-only solve: cameras are variables, the points in x0 are constants."""
        x = np.array(x0, dtype=np.float64, order="C").reshape(-1)
        if x.size != self.n:
            raise ValueError(f"x0 has {x.size} entries, expected {self.n}")
        res = Result()
        fun = np.empty(2 * self.sizes[2]) if want_fun else None
        _check(lib().mmba_solve_pose(self._h, x, C.byref(res), fun.ctypes.data if want_fun else None), self._h)
        return x, res, fun

    def set_x(self, x0):
        _check(lib().mmba_set_x(self._h, self._x(x0)), self._h)

    def solve_resident(self):
        res = Result()
        _check(lib().mmba_solve_resident(self._h, C.byref(res)), self._h)
        return res

    def get_x(self):
        x = np.empty(self.n)
        _check(lib().mmba_get_x(self._h, x), self._h)
        return x

    def log(self):
        n = lib().mmba_get_log(self._h, None, 0)
        rows = (IterLog * max(n, 1))()
        lib().mmba_get_log(self._h, rows, n)
        return [{f: getattr(rows[i], f) for f, _ in IterLog._fields_} for i in range(n)]

    def pcg_history(self):
        """Per inner solve of the last solve: array (iterations + 1, 2) of (||r_k||^2, r_k . Pinv r_k).  Needs an
        engine created with ``profile=2`` (or 3) and the explicit Schur path."""
        out = []
        for i in range(lib().mmba_get_pcg_history(self._h, -1, None, 0)):
            n = lib().mmba_get_pcg_history(self._h, i, None, 0)
            buf = np.empty(max(n, 2))
            lib().mmba_get_pcg_history(self._h, i, buf.ctypes.data, n)
            out.append(buf[:n].reshape(-1, 2))
        return out

    def phase_cycles(self, reset=True):
        out = (C.c_int64 * 64)()
        _check(lib().mmba_get_phase_cycles(self._h, C.byref(out), int(reset)), self._h)
        return np.array(list(out), dtype=np.int64)

    def profile(self):
        launches, ms = (C.c_int64 * K_COUNT)(), (C.c_double * K_COUNT)()
        _check(lib().mmba_get_profile(self._h, C.byref(launches), C.byref(ms)), self._h)
        return {K_NAMES[i]: dict(launches=launches[i], ms=ms[i]) for i in range(K_COUNT)}

    # -- evaluation hooks ---------------------------------------------------------------------
    def _x(self, x):
        x = _c(x, np.float64).reshape(-1)
        if x.size != self.n:
            raise ValueError(f"x has {x.size} entries, expected {self.n}")
        return x

    def residual(self, x):
        f = np.zeros(2 * self.sizes[2])
        _check(lib().mmba_eval_residual(self._h, self._x(x), f), self._h)
        return f

    def jacobian(self, x):
        Jc = np.zeros((self.sizes[2], 2, 6))
        Jp = np.zeros((self.sizes[2], 2, 3))
        _check(lib().mmba_eval_jacobian(self._h, self._x(x), Jc.reshape(-1), Jp.reshape(-1)), self._h)
        return Jc, Jp

    def blocks(self, x):
        nc, npts, _ = self.sizes
        U, V = np.zeros((nc, 6, 6)), np.zeros((npts, 3, 3))
        gc, gp = np.zeros((nc, 6)), np.zeros((npts, 3))
        cost = C.c_double()
        _check(lib().mmba_eval_blocks(self._h, self._x(x), U.reshape(-1), V.reshape(-1), gc.reshape(-1),
                                      gp.reshape(-1), C.byref(cost)), self._h)
        return U, V, gc, gp, cost.value

    def gn_step(self, x, scale, reg):
        p = np.zeros(self.n)
        its, rel = C.c_int64(), C.c_double()
        _check(lib().mmba_eval_gn_step(self._h, self._x(x), self._x(scale), float(reg), p, C.byref(its),
                                       C.byref(rel)), self._h)
        return p, its.value, rel.value

    def reduced_system(self, x, scale, reg):
        """Dense (6 Nc)^2 reduced camera matrix and right-hand side of the explicit Schur path (test hook)."""
        n6 = 6 * self.sizes[0]
        S = np.zeros((n6, n6))
        rhs = np.zeros(n6)
        _check(lib().mmba_eval_reduced_system(self._h, self._x(x), self._x(scale), float(reg), S.reshape(-1), rhs), self._h)
        return S, rhs

    def jnorm2(self, x, s):
        out = C.c_double()
        _check(lib().mmba_eval_jnorm2(self._h, self._x(x), self._x(s), C.byref(out)), self._h)
        return out.value

    def bench_kernel(self, x, kernel_class, iters=20):
        out = C.c_double()
        _check(lib().mmba_bench_kernel(self._h, self._x(x), int(kernel_class), int(iters), C.byref(out)), self._h)
        return out.value


def triangulate(projections, f1, f2, uv1, uv2, device=0, return_ms=False):
    """Two-view DLT triangulation of n tracks at once -> (n,3) points (processor.py:246-261)."""
    proj = _c(projections, np.float64).reshape(-1, 12)
    f1 = _c(f1, np.int64).reshape(-1)
    f2 = _c(f2, np.int64).reshape(-1)
    uv1 = _c(uv1, np.float64).reshape(-1, 2)
    uv2 = _c(uv2, np.float64).reshape(-1, 2)
    n = len(f1)
    if not (len(f2) == len(uv1) == len(uv2) == n):
        raise ValueError("f1, f2, uv1, uv2 must have one entry per track")
    out = np.empty((max(n, 1), 3))
    ms = C.c_double()
    _check(lib().mmba_triangulate(int(device), len(proj), proj.reshape(-1), n, f1 if n else np.zeros(1, np.int64),
                                  f2 if n else np.zeros(1, np.int64), uv1.reshape(-1) if n else np.zeros(2),
                                  uv2.reshape(-1) if n else np.zeros(2), out.reshape(-1), C.byref(ms)))
    out = out[:n]
    return (out, ms.value) if return_ms else out


def rotate(points, rot_vecs, device=0):
    """Row-wise Rodrigues rotation on the GPU (bundleAdjuster.py:7-28)."""
    pts = _c(points, np.float64).reshape(-1, 3)
    rv = _c(rot_vecs, np.float64).reshape(-1, 3)
    if len(pts) != len(rv):
        raise ValueError("points and rot_vecs must have one row per point")
    out = np.empty((max(len(pts), 1), 3))
    _check(lib().mmba_rotate(int(device), len(pts), pts.reshape(-1) if len(pts) else np.zeros(3),
                             rv.reshape(-1) if len(pts) else np.zeros(3), out.reshape(-1)))
    return out[:len(pts)]


def project(points, frame_params, camera_matrix, device=0):
    """Row-wise projection on the GPU (bundleAdjuster.py:31-52)."""
    pts = _c(points, np.float64).reshape(-1, 3)
    fp = _c(frame_params, np.float64)
    fp = fp.reshape(len(pts), -1) if len(pts) else fp.reshape(0, 6)
    if fp.shape[1] < 6:
        raise ValueError("frame_params needs at least 6 columns (rvec | tvec)")
    K = _c(camera_matrix, np.float64).reshape(9)
    out = np.empty((max(len(pts), 1), 2))
    _check(lib().mmba_project(int(device), len(pts), pts.reshape(-1) if len(pts) else np.zeros(3),
                              fp.reshape(-1) if len(pts) else np.zeros(6), fp.shape[1], K, out.reshape(-1)))
    return out[:len(pts)]


# -- host-only helpers (no GPU) -------------------------------------------------------------------
def host_rcm_pattern(n_cams, n_points, cam_idx, pt_idx):
    """Upper-triangle block pattern of the reduced camera matrix: (up_rowptr, up_cols, nnz_full, total_pairs)."""
    cam_idx = _c(cam_idx, np.int64).reshape(-1)
    pt_idx = _c(pt_idx, np.int64).reshape(-1)
    sizes = (C.c_int64 * 3)()
    _check(lib().mmba_host_rcm_pattern(int(n_cams), int(n_points), len(cam_idx), cam_idx, pt_idx, C.byref(sizes), None, None, 0))
    rowptr = np.empty(int(n_cams) + 1, dtype=np.int32)
    cols = np.empty(max(int(sizes[0]), 1), dtype=np.int32)
    _check(lib().mmba_host_rcm_pattern(int(n_cams), int(n_points), len(cam_idx), cam_idx, pt_idx, C.byref(sizes),
                                       rowptr.ctypes.data, cols.ctypes.data, len(cols)))
    return rowptr, cols[:sizes[0]], int(sizes[1]), int(sizes[2])


def host_tr2d(B, g, delta):
    Bc = (C.c_double * 3)(B[0][0], B[0][1], B[1][1])
    gc = (C.c_double * 2)(*g)
    p = (C.c_double * 2)()
    newton = C.c_int()
    _check(lib().mmba_host_tr2d(C.byref(Bc), C.byref(gc), float(delta), C.byref(p), C.byref(newton)))
    return np.array([p[0], p[1]]), bool(newton.value)


def host_min_quadratic_1d(a, b, lb, ub):
    t, y = C.c_double(), C.c_double()
    _check(lib().mmba_host_min_quadratic_1d(a, b, lb, ub, C.byref(t), C.byref(y)))
    return t.value, y.value


def host_update_tr_radius(delta, actual, predicted, step_norm, bound_hit):
    d, r = C.c_double(), C.c_double()
    _check(lib().mmba_host_update_tr_radius(delta, actual, predicted, step_norm, int(bool(bound_hit)),
                                            C.byref(d), C.byref(r)))
    return d.value, r.value


def host_check_termination(dF, F, dx_norm, x_norm, ratio, ftol, xtol):
    code = lib().mmba_host_check_termination(dF, F, dx_norm, x_norm, ratio, ftol, xtol)
    return None if code == 0 else code


def plan(n_cams, n_points, cam_idx, pt_idx, rank=0, nranks=1):
    """Tile plan of one rank as numpy arrays (host only)."""
    cam_idx = _c(cam_idx, np.int64)
    pt_idx = _c(pt_idx, np.int64)
    h = _H()
    _check(lib().mmba_plan_create(C.byref(h), n_cams, n_points, len(cam_idx), cam_idx, pt_idx, rank, nranks))
    try:
        sizes = (C.c_int64 * 8)()
        _check(lib().mmba_plan_sizes(h, C.byref(sizes)))
        keys = ("n_tiles", "n_obs_local", "n_points_local", "point_begin", "point_end", "tile_obs",
                "max_tile_cams", "n_slots")
        out = dict(zip(keys, list(sizes)))
        n_slots = max(out["n_slots"], 1)
        obs_perm = np.full(n_slots, -1, dtype=np.int64)
        point_perm = np.zeros(n_points, dtype=np.int64)
        slot_cam = np.full(n_slots, -1, dtype=np.int32)
        slot_pt = np.full(n_slots, -1, dtype=np.int32)
        _check(lib().mmba_plan_export(h, obs_perm, point_perm, slot_cam, slot_pt))
        out.update(obs_perm=obs_perm[:out["n_slots"]], point_perm=point_perm, slot_cam=slot_cam[:out["n_slots"]],
                   slot_point=slot_pt[:out["n_slots"]])
        nt = out["n_tiles"]
        meta = np.zeros((max(nt, 1), TILE_META_BYTES), dtype=np.uint8)
        tile_cams = np.full((max(nt, 1), 256), -1, dtype=np.int32)
        _check(lib().mmba_plan_raw(h, meta.ctypes.data, tile_cams.ctypes.data))
        out.update(meta=meta[:nt], tile_cams=tile_cams[:nt])
        return out
    finally:
        lib().mmba_plan_destroy(h)
