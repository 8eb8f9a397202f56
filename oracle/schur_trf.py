"""CPU oracle for the engine's solver: scipy's Trust-Region-Reflective outer loop with the LSMR
inner solve replaced by a Schur-complement + block-Jacobi PCG solve.  TEST INFRASTRUCTURE ONLY
(same rules as ``oracle/ba_oracle.py``; nothing under ``meatmodeler_b200/`` imports this).

What it restates
  * outer loop: ``trf_no_bounds`` with ``tr_solver='lsmr'``, ``regularize=True``, ``x_scale='jac'``
    (scipy/optimize/_lsq/trf.py:415-587 — scipy 1.18.1; the reference pins scipy~=1.6.0 and reaches
    it from bundleAdjuster.py:180-192).  The scalar helpers (``solve_trust_region_2d``,
    ``update_tr_radius``, ``check_termination``, ``minimize_quadratic_1d``) are *called from scipy*,
    not re-derived, so the oracle's step rules are the reference's by construction.
  * inner solve: scipy's ``lsmr(J_h, f, damp=sqrt(reg))`` (trf.py:494-495) minimises
    ||J_h p - f||^2 + reg ||p||^2, i.e. solves (J_h^T J_h + reg I) p = J_h^T f.  Here the same
    system is solved by eliminating the 3x3 point blocks and running PCG on the reduced camera
    system with the inverse 6x6 diagonal blocks of the Schur complement as preconditioner,
    zero initial guess, relative-residual stop.
  * Jacobian: analytic blocks (``ba_oracle.jacobian_blocks``) instead of 2-point differences.

Parity status: pinned against the reference's cost trajectory on the committed golden problems
(``tests/golden``): SURVEY.md §0.3 measured per-iteration relative cost differences <= 3.3e-7.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize._lsq.common import (check_termination, minimize_quadratic_1d, solve_lsq_trust_region,
                                        solve_trust_region_2d, update_tr_radius)

from . import ba_oracle as ba


def _segsum(idx, vals, n):
    out = np.empty((n,) + vals.shape[1:], dtype=vals.dtype)
    flat = vals.reshape(len(vals), -1)
    o = out.reshape(n, -1)
    for k in range(flat.shape[1]):
        o[:, k] = np.bincount(idx, weights=flat[:, k], minlength=n)
    return out


class Linearisation:
    """J blocks, residual and normal-equation blocks at one x."""

    def __init__(self, x, K, Nc, Np, fi, pi, uv):
        self.Nc, self.Np, self.fi, self.pi = Nc, Np, fi, pi
        self.r = ba.residuals(x, K, Nc, Np, fi, pi, uv)
        self.Jc, self.Jp = ba.jacobian_blocks(x, K, Nc, Np, fi, pi)
        r2 = self.r.reshape(-1, 2)
        self.U = _segsum(fi, np.einsum("nij,nik->njk", self.Jc, self.Jc), Nc)
        self.V = _segsum(pi, np.einsum("nij,nik->njk", self.Jp, self.Jp), Np)
        self.gc = _segsum(fi, np.einsum("nij,ni->nj", self.Jc, r2), Nc)
        self.gp = _segsum(pi, np.einsum("nij,ni->nj", self.Jp, r2), Np)
        self.cost = 0.5 * float(self.r @ self.r)

    def grad(self):
        return np.hstack((self.gc.ravel(), self.gp.ravel()))

    def colnorm(self):
        """sqrt(column sums of J^2) = sqrt of the block diagonals (common.py:598-610)."""
        return np.sqrt(np.hstack((np.einsum("nii->ni", self.U).ravel(),
                                  np.einsum("nii->ni", self.V).ravel())))

    def jdot(self, s):
        """J s for an unscaled n-vector s -> (No,2)."""
        sc = s[: 6 * self.Nc].reshape(-1, 6)
        sp = s[6 * self.Nc:].reshape(-1, 3)
        return (np.einsum("nij,nj->ni", self.Jc, sc[self.fi]) +
                np.einsum("nij,nj->ni", self.Jp, sp[self.pi]))


def schur_pcg(lin: Linearisation, d, reg, rtol, maxit, atol=0.0, f2=0.0, ktol=0.0):
    """Solve (D J^T J D + reg I) p = D g for p = [p_c | p_p] (scaled variables).

    Returns (p, iterations, relative residual).  All J products use the unscaled blocks with the
    scale folded into the small vectors, exactly as the CUDA kernels do.

    Stopping rules: ||r|| <= rtol ||b||, or (atol > 0) the counterpart of LSMR's second test (lsmr.py:430-459,
    the one that ends scipy's inner solves: ||A^T res|| <= 1e-6 ||A|| ||res||) on the reduced system, where the
    normal-equation residual A^T res is exactly the PCG residual r: ||r|| <= atol ||f||, ||f||^2 = f2.  With
    atol = 1e-7 the iteration counts track the reference's LSMR counts on the golden problems (c1: 18 18 19 14 7 3
    vs 17 22 19 18 13 3) and the cost trajectories stay at the deviation floor of a fully converged solve.

    ktol > 0: the same test with LSMR's growing estimate of ||A|| (the Frobenius norm of the bidiagonal matrix,
    ~ sqrt(k) for unit-norm columns).  LSMR's ||A^T res|| is the residual of a minimal-residual method on the normal
    equations; the minimal-residual norm of a CG process is nu_k with 1/nu_k^2 = sum_{j<=k} 1/||r_j||^2 (residual
    smoothing), so the rule reads  nu_k <= ktol sqrt(k) ||f||.
    """
    Nc, Np, fi, pi, Jc, Jp = lin.Nc, lin.Np, lin.fi, lin.pi, lin.Jc, lin.Jp
    dc = d[: 6 * Nc].reshape(Nc, 6)
    dp = d[6 * Nc:].reshape(Np, 3)
    # M_p = D_p (D_p V D_p + reg I)^-1 D_p  (unscaled-space damped inverse)
    Vh = lin.V * dp[:, :, None] * dp[:, None, :] + reg * np.eye(3)[None]
    M = np.linalg.inv(Vh) * dp[:, :, None] * dp[:, None, :]
    Uh = lin.U * dc[:, :, None] * dc[:, None, :] + reg * np.eye(6)[None]

    def w_apply(zt):
        """sum_i Jc_i^T Jp_i zt[p(i)] per camera (unscaled)."""
        v = np.einsum("nij,nj->ni", Jp, zt[pi])
        return _segsum(fi, np.einsum("nij,ni->nj", Jc, v), Nc)

    def wt_apply(xt):
        """sum_i Jp_i^T Jc_i xt[c(i)] per point (unscaled)."""
        u = np.einsum("nij,nj->ni", Jc, xt[fi])
        return _segsum(pi, np.einsum("nij,ni->nj", Jp, u), Np)

    def matvec(pc):
        t = wt_apply(dc * pc)
        z = np.einsum("nij,nj->ni", M, t)
        return np.einsum("nij,nj->ni", Uh, pc) - dc * w_apply(z)

    # right-hand side and Schur diagonal blocks
    b = dc * (lin.gc - w_apply(np.einsum("nij,nj->ni", M, lin.gp)))
    Wi = np.einsum("nij,nik->njk", Jc, Jp)                      # 6x3 per observation
    Sd = Uh - dc[:, :, None] * dc[:, None, :] * _segsum(
        fi, np.einsum("nij,njk,nlk->nil", Wi, M[pi], Wi), Nc)
    Pinv = np.linalg.inv(Sd)

    x = np.zeros_like(b)
    r = b.copy()
    z = np.einsum("nij,nj->ni", Pinv, r)
    p = z.copy()
    rho = float((r * z).sum())
    bnorm = float(np.sqrt((b * b).sum()))
    it = 0
    rel = 1.0
    inv_nu2 = 1.0 / (bnorm * bnorm) if bnorm > 0 else 0.0
    if bnorm > 0:
        while it < maxit:
            q = matvec(p)
            alpha = rho / float((p * q).sum())
            x += alpha * p
            r -= alpha * q
            it += 1
            rr = float((r * r).sum())
            rel = float(np.sqrt(rr)) / bnorm
            inv_nu2 += 1.0 / rr if rr > 0 else np.inf
            if rel <= rtol or rr <= atol * atol * f2 or 1.0 <= inv_nu2 * ktol * ktol * f2 * it:
                break
            z = np.einsum("nij,nj->ni", Pinv, r)
            rho_new = float((r * z).sum())
            p = z + (rho_new / rho) * p
            rho = rho_new
    # back-substitution: p_p = (V_h + reg)^-1 (D_p g_p - W_h^T p_c)
    dpt = np.einsum("nij,nj->ni", M, lin.gp - wt_apply(dc * x))   # unscaled point step
    pp = dpt / dp
    return np.hstack((x.ravel(), pp.ravel())), it, rel


def solve(x0, K, Nc, Np, fi, pi, uv, ftol=1e-4, xtol=1e-8, gtol=1e-8, max_nfev=None,
          pcg_rtol=1e-10, pcg_maxit=1000, record=None, pcg_atol=1e-7, pcg_ktol=1.23e-6):
    """TRF outer loop (trf.py:415-587) around ``schur_pcg``.  Returns a dict with x, cost, fun,
    nfev, njev, nit, status, optimality and the per-iteration log (cost, reg, Delta, pcg its)."""
    x = np.array(x0, dtype=np.float64)
    lin = Linearisation(x, K, Nc, Np, fi, pi, uv)
    if not np.all(np.isfinite(lin.r)):
        raise ValueError("Residuals are not finite in the initial point.")
    nfev = njev = 1
    cost = lin.cost
    g = lin.grad()
    scale_inv = lin.colnorm()
    scale_inv[scale_inv == 0] = 1
    scale = 1 / scale_inv
    Delta = np.linalg.norm(x * scale_inv)
    if Delta == 0:
        Delta = 1.0
    if max_nfev is None:
        max_nfev = x.size * 100
    status = None
    nit = 0
    log = []
    step_norm = actual = None
    while True:
        g_norm = np.abs(g).max()
        if g_norm < gtol:
            status = 1
        if status is not None or nfev == max_nfev:
            break
        d = scale
        g_h = d * g
        Jg = lin.jdot(d * g_h)                                   # J_h g_h
        a = 0.5 * float((Jg * Jg).sum())
        b = -float(g_h @ g_h)
        to_tr = Delta / np.linalg.norm(g_h)
        ag = minimize_quadratic_1d(a, b, 0, to_tr)[1]
        reg = -ag / Delta ** 2
        gn_h, its, rel = schur_pcg(lin, d, reg, pcg_rtol, pcg_maxit, pcg_atol, 2.0 * cost, pcg_ktol)
        S, _ = np.linalg.qr(np.vstack((g_h, gn_h)).T)
        JS = np.stack((lin.jdot(d * S[:, 0]).ravel(), lin.jdot(d * S[:, 1]).ravel()), axis=1)
        B_S = JS.T @ JS
        g_S = S.T @ g_h
        actual = -1
        while actual <= 0 and nfev < max_nfev:
            p_S, _ = solve_trust_region_2d(B_S, g_S, Delta)
            step_h = S @ p_S
            predicted = -(0.5 * float(p_S @ B_S @ p_S) + float(g_S @ p_S))
            step = d * step_h
            x_new = x + step
            f_new = ba.residuals(x_new, K, Nc, Np, fi, pi, uv)
            nfev += 1
            step_h_norm = np.linalg.norm(step_h)
            if not np.all(np.isfinite(f_new)):
                Delta = 0.25 * step_h_norm
                continue
            cost_new = 0.5 * float(f_new @ f_new)
            actual = cost - cost_new
            Delta_new, ratio = update_tr_radius(Delta, actual, predicted, step_h_norm,
                                                step_h_norm > 0.95 * Delta)
            step_norm = np.linalg.norm(step)
            status = check_termination(actual, cost, step_norm, np.linalg.norm(x), ratio, ftol, xtol)
            if status is not None:
                break
            Delta = Delta_new
        log.append(dict(cost_before=cost, reg=reg, pcg_its=its, pcg_rel=rel, Delta=Delta))
        if actual > 0:
            x = x_new
            cost = cost_new
            lin = Linearisation(x, K, Nc, Np, fi, pi, uv)
            njev += 1
            g = lin.grad()
            scale_inv = np.maximum(lin.colnorm(), scale_inv)
            scale = 1 / scale_inv
        else:
            step_norm = 0
            actual = 0
        nit += 1
        if record is not None:
            record.append(cost)
    if status is None:
        status = 0
    return dict(x=x, cost=cost, fun=lin.r, nfev=nfev, njev=njev, nit=nit, status=status,
                optimality=g_norm, log=log)


def solve_pose(x0, K, Nc, fi, pts_obs, uv, ftol=1e-4, xtol=1e-8, gtol=1e-8, max_nfev=None, record=None):
    """Pose-only TRF with the exact trust-region step (trf.py:415-587 with tr_solver='exact', x_scale=1;
    the reference reaches it from adjustPose, bundleAdjuster.py:232-241) and the analytic camera
    blocks.  ``pts_obs`` (No,3) are the constant 3-D points of the observations.  The dense SVD and
    ``solve_lsq_trust_region`` are scipy's own."""
    n_obs = len(fi)
    pi = np.arange(n_obs)

    def lin(xc):
        x = np.hstack((xc, pts_obs.reshape(-1)))
        r = ba.residuals(x, K, Nc, n_obs, fi, pi, uv)
        Jc, _ = ba.jacobian_blocks(x, K, Nc, n_obs, fi, pi)
        J = np.zeros((2 * n_obs, 6 * Nc))
        rows = np.arange(n_obs)
        for a in range(2):
            for k in range(6):
                J[2 * rows + a, 6 * fi + k] = Jc[:, a, k]
        return r, J

    x = np.array(x0, dtype=np.float64)
    f, J = lin(x)
    if not np.all(np.isfinite(f)):
        raise ValueError("Residuals are not finite in the initial point.")
    m, n = J.shape
    nfev = njev = 1
    cost = 0.5 * float(f @ f)
    g = J.T @ f
    Delta = np.linalg.norm(x)
    if Delta == 0:
        Delta = 1.0
    if max_nfev is None:
        max_nfev = x.size * 100
    alpha = 0.0
    status = None
    nit = 0
    while True:
        g_norm = np.abs(g).max()
        if g_norm < gtol:
            status = 1
        if status is not None or nfev == max_nfev:
            break
        U, s, Vt = np.linalg.svd(J, full_matrices=False)
        V = Vt.T
        uf = U.T @ f
        actual = -1
        while actual <= 0 and nfev < max_nfev:
            step, alpha, _ = solve_lsq_trust_region(n, m, uf, s, V, Delta, initial_alpha=alpha)
            Js = J @ step
            predicted = -(0.5 * float(Js @ Js) + float(g @ step))
            x_new = x + step
            f_new, J_new = lin(x_new)
            nfev += 1
            step_norm = np.linalg.norm(step)
            if not np.all(np.isfinite(f_new)):
                Delta = 0.25 * step_norm
                continue
            cost_new = 0.5 * float(f_new @ f_new)
            actual = cost - cost_new
            Delta_new, ratio = update_tr_radius(Delta, actual, predicted, step_norm, step_norm > 0.95 * Delta)
            status = check_termination(actual, cost, step_norm, np.linalg.norm(x), ratio, ftol, xtol)
            if status is not None:
                break
            alpha *= Delta / Delta_new
            Delta = Delta_new
        if actual > 0:
            x, f, J, cost = x_new, f_new, J_new, cost_new
            g = J.T @ f
            njev += 1
        nit += 1
        if record is not None:
            record.append(cost)
    if status is None:
        status = 0
    return dict(x=x, cost=cost, fun=f, nfev=nfev, njev=njev, nit=nit, status=status, optimality=g_norm)
