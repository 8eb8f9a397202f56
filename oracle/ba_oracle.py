"""CPU oracle for MeatModeler's bundle-adjustment hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy/scipy *restatement* of the reference algorithm; it is the checker the GPU
engine is compared against.  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under ``meatmodeler_b200/`` does.

Parity pinning: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md §4, §8c).  The restatement is therefore pinned against *outputs of the reference itself*,
produced in the build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference/bundleAdjuster.py`` unmodified) and committed under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function here against those vectors.

The optimiser arithmetic of the reference lives in a third-party dependency that is not under
``/root/reference``: scipy (``requirements.txt:8`` pins ``scipy~=1.6.0``; this image has 1.18.1).
``solve_reference_path`` therefore calls ``scipy.optimize.least_squares`` with the exact keyword
arguments of the reference call site (``bundleAdjuster.py:180-192``), which is what the reference
does; ``oracle/schur_trf.py`` restates scipy's published TRF algorithm for the engine's own solver.

Every function cites the reference lines it follows.  All functions are dtype-generic: they run
unchanged in ``numpy.longdouble`` (80-bit), which the Jacobian parity tests rely on.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------------------------
# residual model  (bundleAdjuster.py:7-52, 81-102)
# ----------------------------------------------------------------------------------------------

def rotate(points, rot_vecs):
    """Rodrigues rotation of ``points`` (N,3) by axis-angle vectors ``rot_vecs`` (N,3).

    Follows bundleAdjuster.py:7-28: theta = |w|, v = w/theta with 0/0 -> 0, result =
    cos(theta) X + sin(theta) (v x X) + (v.X)(1-cos(theta)) v, so theta == 0 returns X exactly.
    """
    angle = np.sqrt((rot_vecs * rot_vecs).sum(axis=1, keepdims=True))
    with np.errstate(invalid="ignore", divide="ignore"):
        axis = np.where(angle > 0, rot_vecs / angle, np.zeros_like(rot_vecs))
    c = np.cos(angle)
    s = np.sin(angle)
    along = (points * axis).sum(axis=1, keepdims=True)
    return c * points + s * np.cross(axis, points) + along * (1 - c) * axis


def project(points, frame_params, camera_matrix):
    """Pinhole projection with one shared intrinsic matrix (bundleAdjuster.py:31-52).

    Xc = R(w) X + t ; q = K Xc ; (u, v) = (q0/q2, q1/q2).  ``frame_params`` rows are [w | t].
    """
    cam = rotate(points, frame_params[:, :3]) + frame_params[:, 3:6]
    q = cam @ np.asarray(camera_matrix, dtype=cam.dtype).T
    return q[:, :2] / q[:, 2:3]


def residuals(x, camera_matrix, n_frames, n_points, frame_indices, point_indices, points_2d):
    """Residual vector f(x), interleaved (du0, dv0, du1, ...)  (bundleAdjuster.py:81-102).

    ``x`` = [6*n_frames camera parameters | 3*n_points coordinates] (bundleAdjuster.py:175-176).
    """
    cams = x[: 6 * n_frames].reshape(n_frames, 6)
    pts = x[6 * n_frames:].reshape(n_points, 3)
    uv = project(pts[point_indices], cams[frame_indices], camera_matrix)
    return (uv - points_2d).ravel()


def sparsity(n_frames, n_points, frame_indices, point_indices):
    """Structural Jacobian pattern as the reference builds it (bundleAdjuster.py:55-78).

    Rows 2i and 2i+1 carry ones in the 6 columns of camera fi[i] and the 3 columns of point pi[i].
    Built through ``lil_matrix`` fancy assignment exactly like the reference so that the CPU
    baseline pays the same construction cost.
    """
    from scipy.sparse import lil_matrix

    n_obs = frame_indices.size
    pattern = lil_matrix((2 * n_obs, 6 * n_frames + 3 * n_points), dtype=int)
    rows = np.arange(n_obs)
    for k in range(6):
        for parity in (0, 1):
            pattern[2 * rows + parity, 6 * frame_indices + k] = 1
    for k in range(3):
        for parity in (0, 1):
            pattern[2 * rows + parity, 6 * n_frames + 3 * point_indices + k] = 1
    return pattern


def frame_parameters(extrinsics):
    """(Nc,3|4,4) extrinsic matrices -> flat [w0,t0,w1,t1,...]  (bundleAdjuster.py:105-134).

    theta = arccos((tr R - 1)/2); axis from the skew part / (2 sin theta) with 0/0 -> 0.
    """
    ext = np.asarray(extrinsics)
    rot = ext[:, :3, :3]
    angle = np.arccos((rot[:, 0, 0] + rot[:, 1, 1] + rot[:, 2, 2] - 1) / 2)
    denom = 2 * np.sin(angle)
    skew = np.stack((rot[:, 2, 1] - rot[:, 1, 2],
                     rot[:, 0, 2] - rot[:, 2, 0],
                     rot[:, 1, 0] - rot[:, 0, 1]), axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        axis = np.nan_to_num(skew / denom[:, None])
    return np.hstack((axis * angle[:, None], ext[:, :3, 3])).reshape(-1)


def rodrigues_matrix(rvec):
    """3x3 rotation of one axis-angle vector (what cv2.Rodrigues returns, bundleAdjuster.py:153)."""
    rvec = np.asarray(rvec, dtype=np.float64).reshape(3)
    angle = np.linalg.norm(rvec)
    if angle == 0:
        return np.eye(3)
    k = rvec / angle
    kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(angle) * np.eye(3) + np.sin(angle) * kx + (1 - np.cos(angle)) * np.outer(k, k)


def reformat_point_result(x, n_frames, n_points):
    """x -> ((Np,3) points, list of Nc 4x4 extrinsics)  (bundleAdjuster.py:137-157)."""
    cams = x[: 6 * n_frames].reshape(n_frames, 6)
    pts = x[6 * n_frames:].reshape(n_points, 3)
    out = []
    for row in cams:
        m = np.eye(4)
        m[:3, :3] = rodrigues_matrix(row[:3])
        m[:3, 3] = row[3:]
        out.append(m)
    return pts, out


# ----------------------------------------------------------------------------------------------
# the reference solve: scipy least_squares with the reference's kwargs (bundleAdjuster.py:160-194)
# ----------------------------------------------------------------------------------------------

def solve_reference_path(extrinsics, camera_matrix, points_3d, points_2d, frame_indices,
                         point_indices, verbose=0, max_nfev=None, record=None):
    """CPU restatement of ``adjustPoints`` (bundleAdjuster.py:160-194): pack, sparsity pattern,
    ``least_squares(method='trf', jac_sparsity=A, x_scale='jac', ftol=1e-4)`` with 2-point finite
    differences, unpack.  Returns the scipy ``OptimizeResult``; ``record`` (a list) receives the
    cost after every outer iteration (scipy passes it through ``callback=``, whose parameter must
    be called ``intermediate_result``).
    """
    from scipy.optimize import least_squares

    ext = np.asarray(extrinsics, dtype=np.float64)
    pts = np.asarray(points_3d, dtype=np.float64)
    n_frames, n_points = len(ext), len(pts)
    x0 = np.hstack((frame_parameters(ext), pts.reshape(n_points * 3)))
    pattern = sparsity(n_frames, n_points, frame_indices, point_indices)

    def on_iter(intermediate_result):
        if record is not None:
            record.append(float(intermediate_result.cost))

    return least_squares(residuals, x0, jac_sparsity=pattern, verbose=verbose, x_scale="jac",
                         ftol=1e-4, method="trf", max_nfev=max_nfev, callback=on_iter,
                         args=(camera_matrix, n_frames, n_points, frame_indices, point_indices,
                               points_2d))


def adjust_points(extrinsics, camera_matrix, points_3d, points_2d, frame_indices, point_indices,
                  verbose=0):
    """Same signature and return value as the reference ``adjustPoints``."""
    res = solve_reference_path(extrinsics, camera_matrix, points_3d, points_2d, frame_indices,
                               point_indices, verbose=verbose)
    return reformat_point_result(res.x, len(extrinsics), len(points_3d))


# ----------------------------------------------------------------------------------------------
# the pose-only path (bundleAdjuster.py:197-243)
# ----------------------------------------------------------------------------------------------

def board_points(pattern_size):
    """Chessboard of adjustPose: 4x3 grid times 2 in the x-z plane, float32 (bundleAdjuster.py:220-223)."""
    pts = np.zeros((pattern_size, 3), np.float32)
    grid = np.mgrid[0:4, 0:3].T.reshape(-1, 2) * 2
    pts[:, 0] = grid[:, 0]
    pts[:, 2] = grid[:, 1]
    return pts


def pose_residuals(x, camera_matrix, n_frames, frame_indices, point_indices, points_3d, points_2d):
    """poseFun (bundleAdjuster.py:206-211)."""
    cams = x.reshape(n_frames, 6)
    uv = project(points_3d[point_indices], cams[frame_indices], camera_matrix)
    return (uv - points_2d).ravel()


def solve_pose_reference_path(extrinsics, camera_matrix, points_2d, verbose=0, record=None):
    """CPU restatement of ``adjustPose`` (bundleAdjuster.py:214-243): dense
    ``least_squares(poseFun, parameters, ftol=1e-4)`` -> trf, 2-point dense finite differences,
    tr_solver='exact', x_scale=1.  Returns the scipy result."""
    from scipy.optimize import least_squares

    ext = np.asarray(extrinsics, dtype=np.float64)
    n_frames = len(ext)
    pattern = int(len(points_2d) / n_frames)
    pts = board_points(pattern)
    fi = np.repeat(np.arange(n_frames), pattern)
    pi = np.repeat([np.arange(pattern)], n_frames, axis=0).reshape(pattern * n_frames)

    def on_iter(intermediate_result):
        if record is not None:
            record.append(float(intermediate_result.cost))

    return least_squares(pose_residuals, frame_parameters(ext), verbose=verbose, ftol=1e-4, callback=on_iter,
                         args=(camera_matrix, n_frames, fi, pi, pts, points_2d))


# ----------------------------------------------------------------------------------------------
# Jacobian blocks
# ----------------------------------------------------------------------------------------------

def _skew(v):
    z = np.zeros_like(v[..., 0])
    return np.stack((np.stack((z, -v[..., 2], v[..., 1]), -1),
                     np.stack((v[..., 2], z, -v[..., 0]), -1),
                     np.stack((-v[..., 1], v[..., 0], z), -1)), -2)


def rotation_and_tangent(rvecs):
    """Per-camera R(w) and Q(w) = R(w) (a I - b [w]x + c w w^T) with a = sin t/t,
    b = (1-cos t)/t^2, c = (t - sin t)/t^3, so that d(R X)/dw = -[R X]x Q.

    Equivalent to SURVEY.md §8a's closed form -R [X]x (w w^T + (R^T - I)[w]x)/t^2; the
    coefficients are evaluated by their Taylor series below t = 0.05 so t -> 0 is regular.
    """
    w = np.asarray(rvecs)
    dt = w.dtype
    t2 = (w * w).sum(axis=1)
    t = np.sqrt(t2)
    small = t < 0.05
    ts = np.where(small, np.ones_like(t), t)
    half = np.sin(ts / 2)
    a_big = np.sin(ts) / ts
    b_big = 2 * half * half / (ts * ts)
    c_big = (ts - np.sin(ts)) / (ts * ts * ts)
    one = dt.type(1)
    a_ser = one - t2 / 6 * (one - t2 / 20 * (one - t2 / 42 * (one - t2 / 72)))
    b_ser = one / 2 * (one - t2 / 12 * (one - t2 / 30 * (one - t2 / 56 * (one - t2 / 90))))
    c_ser = one / 6 * (one - t2 / 20 * (one - t2 / 42 * (one - t2 / 72 * (one - t2 / 110))))
    a = np.where(small, a_ser, a_big)[:, None, None]
    b = np.where(small, b_ser, b_big)[:, None, None]
    c = np.where(small, c_ser, c_big)[:, None, None]
    eye = np.eye(3, dtype=dt)[None]
    wx = _skew(w)
    wwt = w[:, :, None] * w[:, None, :]
    rot = eye + a * wx + b * (wwt - t2[:, None, None] * eye)
    tangent = a * eye - b * wx + c * wwt
    return rot, rot @ tangent


def jacobian_blocks(x, camera_matrix, n_frames, n_points, frame_indices, point_indices):
    """Analytic Jacobian blocks of ``residuals``: (No,2,6) camera blocks with column order
    (w0,w1,w2,t0,t1,t2) and (No,2,3) point blocks  (closed form of SURVEY.md §8a; the reference
    obtains the same matrix by finite differences, scipy/optimize/_numdiff.py:770-893).
    """
    cams = x[: 6 * n_frames].reshape(n_frames, 6)
    pts = x[6 * n_frames:].reshape(n_points, 3)
    K = np.asarray(camera_matrix, dtype=x.dtype)
    rot, q_mat = rotation_and_tangent(cams[:, :3])
    R = rot[frame_indices]
    Q = q_mat[frame_indices]
    X = pts[point_indices]
    Y = np.einsum("nij,nj->ni", R, X)
    Xc = Y + cams[frame_indices, 3:6]
    q = Xc @ K.T
    inv = 1 / q[:, 2]
    u = q[:, 0] * inv
    v = q[:, 1] * inv
    A = np.empty((len(X), 2, 3), dtype=x.dtype)
    A[:, 0, :] = (K[0][None, :] - u[:, None] * K[2][None, :]) * inv[:, None]
    A[:, 1, :] = (K[1][None, :] - v[:, None] * K[2][None, :]) * inv[:, None]
    Jc = np.empty((len(X), 2, 6), dtype=x.dtype)
    Jc[:, :, 3:] = A
    Jc[:, :, :3] = -np.einsum("nij,njk,nkl->nil", A, _skew(Y), Q)
    Jp = np.einsum("nij,njk->nik", A, R)
    return Jc, Jp


def jacobian_blocks_fd(x, camera_matrix, n_frames, n_points, frame_indices, point_indices,
                       rel_step=1e-6):
    """Central differences of ``project`` per (camera, point) pair in ``numpy.longdouble``
    (h = 1e-6*max(1,|x|) gives ~7e-13 self-consistency).  This is the
    Jacobian oracle; scipy's own 2-point matrix is only ~5e-8 accurate.
    """
    ld = np.longdouble
    xl = np.asarray(x, dtype=ld)
    K = np.asarray(camera_matrix, dtype=ld)
    cams = xl[: 6 * n_frames].reshape(n_frames, 6)[frame_indices]
    pts = xl[6 * n_frames:].reshape(n_points, 3)[point_indices]
    par = np.hstack((cams, pts))
    n_obs = len(par)
    out = np.empty((n_obs, 2, 9), dtype=ld)
    for k in range(9):
        h = ld(rel_step) * np.maximum(ld(1), np.abs(par[:, k]))
        hi = par.copy()
        lo = par.copy()
        hi[:, k] += h
        lo[:, k] -= h
        d = project(hi[:, 6:], hi[:, :6], K) - project(lo[:, 6:], lo[:, :6], K)
        out[:, :, k] = d / (2 * h)[:, None]
    return out[:, :, :6], out[:, :, 6:]


def block_sums(Jc, Jp, r, n_frames, n_points, frame_indices, point_indices):
    """Normal-equation blocks from per-observation blocks: U (Nc,6,6) = sum Jc^T Jc,
    V (Np,3,3) = sum Jp^T Jp, gc (Nc,6) = sum Jc^T r, gp (Np,3) = sum Jp^T r
    (what J^T J and J^T f of scipy/optimize/_lsq/common.py:590-610 contain block-wise).
    """
    r2 = r.reshape(-1, 2)
    U = np.zeros((n_frames, 6, 6), dtype=Jc.dtype)
    V = np.zeros((n_points, 3, 3), dtype=Jc.dtype)
    gc = np.zeros((n_frames, 6), dtype=Jc.dtype)
    gp = np.zeros((n_points, 3), dtype=Jc.dtype)
    np.add.at(U, frame_indices, np.einsum("nij,nik->njk", Jc, Jc))
    np.add.at(V, point_indices, np.einsum("nij,nik->njk", Jp, Jp))
    np.add.at(gc, frame_indices, np.einsum("nij,ni->nj", Jc, r2))
    np.add.at(gp, point_indices, np.einsum("nij,ni->nj", Jp, r2))
    return U, V, gc, gp
