"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, numpy, scipy, cv2):
    python tests/golden/make_golden.py
The reference module is imported by path, never copied.  The .npz files are committed; nothing at
test time reads /root/reference.

Files
  small.npz   Nc=6 (camera 0 has rvec == 0), Np=40, No=214: inputs, reference ``pointFun`` in float64
              and longdouble, ``frameParameters``, longdouble central-difference Jacobian blocks of
              the reference ``project``, ``pointAdjustmentSparsity`` pattern, and the reference
              ``adjustPoints`` outputs + scipy per-iteration costs.
  c1.npz      BASELINE configs[0] (20 / 2 000 / 40 000, seed 1, hard init): reference cost
              trajectory, nfev, final cost / optimality, final points checksum, LSMR iterations.
  mid.npz     Nc=60, Np=1500, No=9000 sparse-visibility problem: same scalars.
  pose.npz    adjustPose (pose-only, dense least_squares) on a 9-frame synthetic chessboard sequence:
              inputs, poseFun at x0, per-iteration costs, final parameters and returned 3x4 matrices.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference")

import bundleAdjuster as ref  # noqa: E402  (the reference, unmodified)
import scipy  # noqa: E402
from scipy.optimize import least_squares  # noqa: E402

from meatmodeler_b200 import synth  # noqa: E402


def reference_trajectory(prob):
    """The reference call (bundleAdjuster.py:180-192) with a cost-recording callback."""
    ext, K, pts, uv, fi, pi = prob.args()
    nc, npts = len(ext), len(pts)
    x0 = np.hstack((ref.frameParameters(ext), pts.reshape(npts * 3)))
    A = ref.pointAdjustmentSparsity(nc, npts, fi, pi)
    costs = []
    import scipy.optimize._lsq.trf as T
    lsmr_its = []
    orig = T.lsmr

    def spy(*a, **k):
        out = orig(*a, **k)
        lsmr_its.append(int(out[2]))
        return out

    T.lsmr = spy
    try:
        res = least_squares(ref.pointFun, x0, jac_sparsity=A, verbose=0, x_scale="jac", ftol=1e-4, method="trf",
                            args=(K, nc, npts, fi, pi, uv),
                            callback=lambda intermediate_result: costs.append(float(intermediate_result.cost)))
    finally:
        T.lsmr = orig
    f0 = ref.pointFun(x0, K, nc, npts, fi, pi, uv)
    return x0, res, np.array([0.5 * f0 @ f0] + costs), np.array(lsmr_its)


def jac_blocks_ld(x, K, nc, npts, fi, pi, rel=1e-6):
    """Longdouble central differences of the reference ``project`` per (camera, point) pair."""
    ld = np.longdouble
    xl = x.astype(ld)
    cams = xl[:6 * nc].reshape(nc, 6)[fi]
    pts = xl[6 * nc:].reshape(npts, 3)[pi]
    par = np.hstack((cams, pts))
    Kl = K.astype(ld)
    out = np.empty((len(par), 2, 9), dtype=ld)
    for k in range(9):
        h = ld(rel) * np.maximum(ld(1), np.abs(par[:, k]))
        hi, lo = par.copy(), par.copy()
        hi[:, k] += h
        lo[:, k] -= h
        d = ref.project(hi[:, 6:], hi[:, :6], Kl) - ref.project(lo[:, 6:], lo[:, :6], Kl)
        out[:, :, k] = d / (2 * h)[:, None]
    return out[:, :, :6].astype(np.float64), out[:, :, 6:].astype(np.float64)


def small_problem():
    rng = np.random.default_rng(11)
    nc, npts = 6, 40
    K = np.array([[900.0, 0.5, 310.0], [0, 880.0, 250.0], [0, 0, 1.0]])
    # camera 0 is the world frame (rvec == 0 exactly); the others look roughly down +z
    ext = np.zeros((nc, 3, 4))
    for c in range(nc):
        w = np.zeros(3) if c == 0 else rng.normal(0, 0.25, 3)
        th = np.linalg.norm(w)
        if th == 0:
            R = np.eye(3)
        else:
            k = w / th
            Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
            R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        ext[c, :, :3] = R
        ext[c, :, 3] = rng.normal(0, 0.3, 3) + np.array([0, 0, 6.0])
    X = rng.normal(0, 1, (npts, 3))
    lengths = rng.integers(2, nc + 1, npts)
    fi = np.concatenate([np.sort(rng.choice(nc, L, replace=False)) for L in lengths]).astype(np.int64)
    pi = np.repeat(np.arange(npts, dtype=np.int64), lengths)
    params = ref.frameParameters(ext).reshape(nc, 6)
    uv = ref.project(X[pi], params[fi], K) + rng.normal(0, 0.5, (len(fi), 2))
    # shuffle the observation order: the signature accepts any order
    perm = rng.permutation(len(fi))
    fi, pi, uv = fi[perm], pi[perm], uv[perm]
    X0 = X + rng.normal(0, 0.05, X.shape)
    ext0 = ext.copy()
    ext0[1:, :, 3] += rng.normal(0, 0.03, (nc - 1, 3))
    return synth.Problem(ext0, K, X0.reshape(npts, 1, 3), uv, fi, pi, ext, X)


def main():
    # ---- small.npz -------------------------------------------------------------------------
    prob = small_problem()
    ext, K, pts, uv, fi, pi = prob.args()
    nc, npts = len(ext), len(pts)
    x0, res, costs, lsmr_its = reference_trajectory(prob)
    f64 = ref.pointFun(x0, K, nc, npts, fi, pi, uv)
    fld = ref.pointFun(x0.astype(np.longdouble), K, nc, npts, fi, pi, uv)
    Jc, Jp = jac_blocks_ld(x0, K, nc, npts, fi, pi)
    pattern = ref.pointAdjustmentSparsity(nc, npts, fi, pi).tocsr()
    with contextlib.redirect_stdout(io.StringIO()):
        adj_pts, adj_ext = ref.adjustPoints(ext, K, pts, uv, fi, pi)
    # a second evaluation point away from x0 (all cameras rotated, camera 0 included)
    rng = np.random.default_rng(5)
    x1 = x0 + rng.normal(0, 0.02, x0.shape)
    f64_1 = ref.pointFun(x1, K, nc, npts, fi, pi, uv)
    Jc1, Jp1 = jac_blocks_ld(x1, K, nc, npts, fi, pi)
    np.savez_compressed(
        os.path.join(HERE, "small.npz"), ext=ext, K=K, pts=pts, uv=uv, fi=fi, pi=pi, x0=x0,
        frame_parameters=ref.frameParameters(ext), f64=f64, fld=fld.astype(np.float64),
        fld_lo=(fld - fld.astype(np.float64).astype(np.longdouble)).astype(np.float64), Jc=Jc, Jp=Jp,
        x1=x1, f64_1=f64_1, Jc1=Jc1, Jp1=Jp1,
        pattern_indices=pattern.indices.astype(np.int32), pattern_indptr=pattern.indptr.astype(np.int64),
        ref_x=res.x, ref_cost=res.cost, ref_costs=costs, ref_nfev=res.nfev, ref_njev=res.njev, ref_status=res.status,
        ref_optimality=res.optimality, ref_fun=res.fun, ref_lsmr_its=lsmr_its, adj_points=adj_pts,
        adj_extrinsics=np.array(adj_ext),
        versions=np.array([np.__version__, scipy.__version__]))
    print("small:", costs, res.nfev, res.status, lsmr_its)

    # ---- c1.npz / mid.npz ------------------------------------------------------------------
    for name, prob in (("c1", synth.make_config("C1", hard=True)),
                       ("mid", synth.make_problem(60, 1500, 9000, seed=7, hard=True, windowed=False))):
        ext, K, pts, uv, fi, pi = prob.args()
        x0, res, costs, lsmr_its = reference_trajectory(prob)
        npts = len(pts)
        rms = np.sqrt(np.mean(np.sum(res.fun.reshape(-1, 2) ** 2, axis=1)))
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), sizes=np.array(prob.sizes), x0_checksum=float(np.sum(x0)),
            uv_checksum=float(np.sum(uv)), ref_costs=costs, ref_cost=res.cost, ref_nfev=res.nfev, ref_njev=res.njev,
            ref_status=res.status, ref_optimality=res.optimality, ref_rms=rms, ref_lsmr_its=lsmr_its,
            ref_points_sample=res.x[6 * len(ext):].reshape(npts, 3)[:: max(1, npts // 50)],
            ref_x_cams=res.x[:6 * len(ext)], versions=np.array([np.__version__, scipy.__version__]))
        print(name, costs, res.nfev, res.status, rms, lsmr_its)


def pose_golden():
    """adjustPose of the unmodified reference on a synthetic chessboard sequence -> pose.npz."""
    rng = np.random.default_rng(23)
    n_frames = 9
    K = np.array([[950.0, 0, 320.0], [0, 940.0, 240.0], [0, 0, 1.0]])
    board = np.zeros((12, 3), np.float32)
    grid = np.mgrid[0:4, 0:3].T.reshape(-1, 2) * 2
    board[:, 0] = grid[:, 0]
    board[:, 2] = grid[:, 1]
    ext = np.zeros((n_frames, 3, 4))
    for c in range(n_frames):
        w = rng.normal(0, 0.3, 3) + np.array([0.9, 0.0, 0.0])     # look down at the x-z plane
        th = np.linalg.norm(w)
        k = w / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        ext[c, :, :3] = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        ext[c, :, 3] = np.array([-3.0, -1.0, 14.0]) + rng.normal(0, 0.8, 3)
    params = ref.frameParameters(ext).reshape(n_frames, 6)
    fi = np.repeat(np.arange(n_frames), 12)
    pi = np.tile(np.arange(12), n_frames)
    uv = ref.project(board[pi].astype(np.float64), params[fi], K) + rng.normal(0, 0.3, (len(fi), 2))
    ext0 = ext.copy()
    ext0[:, :, 3] += rng.normal(0, 0.4, (n_frames, 3))
    costs = []
    x0 = ref.frameParameters(ext0)
    res = least_squares(ref.poseFun, x0, verbose=0, ftol=1e-4,
                        args=(K, n_frames, fi, pi, board, uv),
                        callback=lambda intermediate_result: costs.append(float(intermediate_result.cost)))
    f0 = ref.poseFun(x0, K, n_frames, fi, pi, board, uv)
    with contextlib.redirect_stdout(io.StringIO()):
        out = ref.adjustPose(ext0, K, uv)
    np.savez_compressed(os.path.join(HERE, "pose.npz"), ext0=ext0, K=K, uv=uv, x0=x0, f0=f0,
                        ref_costs=np.array([0.5 * f0 @ f0] + costs), ref_x=res.x, ref_cost=res.cost, ref_nfev=res.nfev,
                        ref_status=res.status, ref_optimality=res.optimality, adj_extrinsics=np.array(out),
                        versions=np.array([np.__version__, scipy.__version__]))
    print("pose:", np.array([0.5 * f0 @ f0] + costs), res.nfev, res.status)


if __name__ == "__main__":
    main()
    pose_golden()
