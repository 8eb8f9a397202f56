"""CPU prototype (scipy.sparse) of stronger preconditioners for the reduced camera system on a long camera chain.

    python tools/precond_prototype.py [scale]        # C4 shape scaled down (all 1 778 cameras kept), default 0.05

Builds the scaled, damped reduced camera matrix S = U - W V'^-1 W^T from the oracle's Jacobian blocks at the
(hard) starting point and counts PCG iterations to a relative residual of 1e-8 for
  jacobi6        6x6 block-Jacobi (what rcm_pcg_kernel applies today),
  block13        exact inverse of the diagonal block of 13 consecutive cameras (one PCG CTA's range),
  *+coarseN      additive two-level: the smoother + Z A_c^-1 Z^T with piecewise-constant aggregates of N cameras
                 (6 coarse unknowns per aggregate, A_c = Z^T S Z).
Test infrastructure / design evidence only (round-1 result, 1 778 cameras / 250 k observations):
  reg = 1e-4:  jacobi6 554 | block13 247 | jacobi6+coarse13 128 | jacobi6+coarse26 143 | jacobi6+coarse52 228 | block13+coarse13 101
  reg = 1e-6:  jacobi6 >3000 | block13 1660 | jacobi6+coarse13 568 | jacobi6+coarse26 611 | jacobi6+coarse52 949 | block13+coarse13 469
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from meatmodeler_b200 import synth  # noqa: E402
from oracle import ba_oracle as ba  # noqa: E402
from oracle import schur_trf as st  # noqa: E402


def reduced_system(lin, d, reg):
    no, nc, npts, fi, pi = len(lin.fi), lin.Nc, lin.Np, lin.fi, lin.pi
    rows = np.repeat(np.arange(2 * no), 6)
    cols = (6 * fi[:, None, None] + np.arange(6)[None, None, :] + np.zeros((1, 2, 1), dtype=np.int64)).reshape(-1)
    Jc = sp.csr_matrix((lin.Jc.reshape(-1), (rows, cols)), shape=(2 * no, 6 * nc)) @ sp.diags(d[:6 * nc])
    rows = np.repeat(np.arange(2 * no), 3)
    cols = (3 * pi[:, None, None] + np.arange(3)[None, None, :] + np.zeros((1, 2, 1), dtype=np.int64)).reshape(-1)
    Jp = sp.csr_matrix((lin.Jp.reshape(-1), (rows, cols)), shape=(2 * no, 3 * npts)) @ sp.diags(d[6 * nc:])
    U, W = (Jc.T @ Jc).tocsr(), (Jc.T @ Jp).tocsr()
    V = (Jp.T @ Jp + reg * sp.identity(3 * npts)).tocoo()
    Vb = np.zeros((npts, 3, 3))
    Vb[V.row // 3, V.row % 3, V.col % 3] = V.data
    r_ = np.repeat(np.arange(3 * npts), 3)
    c_ = (3 * (np.arange(3 * npts) // 3)[:, None] + np.arange(3)[None, :]).reshape(-1)
    Vinv = sp.csr_matrix((np.linalg.inv(Vb).reshape(-1), (r_, c_)), shape=(3 * npts, 3 * npts))
    g = d * lin.grad()
    return (U + reg * sp.identity(6 * nc) - W @ Vinv @ W.T).tocsr(), g[:6 * nc] - W @ (Vinv @ g[6 * nc:])


def pcg(S, b, apply_m, tol=1e-8, maxit=3000):
    x, r = np.zeros_like(b), b.copy()
    z = apply_m(r)
    p, rho, b2 = z.copy(), r @ z, b @ b
    for it in range(1, maxit + 1):
        q = S @ p
        a = rho / (p @ q)
        x += a * p
        r -= a * q
        if r @ r <= tol * tol * b2:
            break
        z = apply_m(r)
        rho_new = r @ z
        p = z + (rho_new / rho) * p
        rho = rho_new
    return it


def block_jacobi(S, nc, cams_per_block):
    n = 6 * nc
    cuts = [(6 * c0, min(6 * (c0 + cams_per_block), n)) for c0 in range(0, nc, cams_per_block)]
    invs = [np.linalg.inv(S[i0:i1, i0:i1].toarray()) for i0, i1 in cuts]

    def apply(r):
        z = np.empty_like(r)
        for (i0, i1), m in zip(cuts, invs):
            z[i0:i1] = m @ r[i0:i1]
        return z
    return apply


def coarse(S, nc, cams_per_aggregate):
    n = 6 * nc
    idx = np.arange(n)
    Z = sp.csr_matrix((np.ones(n), (idx, 6 * ((idx // 6) // cams_per_aggregate) + idx % 6)))
    Ainv = np.linalg.inv((Z.T @ S @ Z).toarray())
    return lambda r: Z @ (Ainv @ (Z.T @ r))


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
    prob = synth.make_config("C4", hard=True, scale=scale)
    ext, K, pts, uv, fi, pi = prob.args()
    nc, npts = len(ext), len(pts)
    x0 = np.hstack((ba.frame_parameters(ext), np.asarray(pts).reshape(-1)))
    lin = st.Linearisation(x0, K, nc, npts, fi, pi, uv)
    d = 1.0 / np.where(lin.colnorm() == 0, 1.0, lin.colnorm())
    print(f"{nc} cameras, {npts} points, {len(fi)} observations")
    for reg in (1e-4, 1e-6):
        S, b = reduced_system(lin, d, reg)
        j6, b13 = block_jacobi(S, nc, 1), block_jacobi(S, nc, 13)
        out = {"jacobi6": pcg(S, b, j6), "block13": pcg(S, b, b13)}
        for agg in (13, 26, 52):
            c = coarse(S, nc, agg)
            out[f"jacobi6+coarse{agg}"] = pcg(S, b, lambda r: j6(r) + c(r))
        c13 = coarse(S, nc, 13)
        out["block13+coarse13"] = pcg(S, b, lambda r: b13(r) + c13(r))
        print(f"reg {reg:g}: {out}", flush=True)


if __name__ == "__main__":
    main()
