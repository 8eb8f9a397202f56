"""Synthetic turntable bundle-adjustment problems (SURVEY.md §8d recipe).

Inputs are produced exactly in the shapes ``processor.py:465-470`` hands to ``adjustPoints``:
extrinsics (Nc,3,4) f64, K (3,3) f64, points (Np,1,3) f64, observations (No,2) f64 and int64
index arrays, observations grouped by point (``processor.py:280-289``).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

CONFIGS = {
    # name: (Nc, Np, No)   — BASELINE.json configs[0..4]
    "C1": (20, 2_000, 40_000),
    "C2": (200, 50_000, 1_000_000),
    "C3": (1_723, 156_000, 679_000),
    "C4": (1_778, 993_000, 5_000_000),
    "C5": (10_000, 2_000_000, 20_000_000),
}
CONFIG_SEEDS = {"C1": 1, "C2": 2, "C3": 3, "C4": 4, "C5": 5}


@dataclass
class Problem:
    extrinsics: np.ndarray      # (Nc,3,4) initial guess
    K: np.ndarray               # (3,3)
    points: np.ndarray          # (Np,1,3) initial guess
    uv: np.ndarray              # (No,2)
    cam_idx: np.ndarray         # (No,) int64
    pt_idx: np.ndarray          # (No,) int64
    true_extrinsics: np.ndarray
    true_points: np.ndarray

    @property
    def sizes(self):
        return len(self.extrinsics), len(self.points), len(self.uv)

    def args(self):
        """Positional arguments of ``adjustPoints`` (bundleAdjuster.py:160)."""
        return (self.extrinsics, self.K, self.points, self.uv, self.cam_idx, self.pt_idx)


def ring_cameras(n_cams: int) -> np.ndarray:
    """Cameras on a ring of radius 8 at height 3 looking at the origin; returns (Nc,3,4) [R|t]."""
    a = 2 * np.pi * np.arange(n_cams) / n_cams
    C = np.stack((8 * np.cos(a), np.full(n_cams, 3.0), 8 * np.sin(a)), axis=1)
    z = -C / np.linalg.norm(C, axis=1, keepdims=True)
    up = np.array([0.0, 1.0, 0.0])
    xax = np.cross(up[None, :], z)
    xax /= np.linalg.norm(xax, axis=1, keepdims=True)
    yax = np.cross(z, xax)
    R = np.stack((xax, yax, z), axis=1)
    t = -np.einsum("nij,nj->ni", R, C)
    return np.concatenate((R, t[:, :, None]), axis=2)


def _project(ext, K, X, cam_idx, pt_idx):
    Xc = np.einsum("nij,nj->ni", ext[cam_idx, :, :3], X[pt_idx]) + ext[cam_idx, :, 3]
    q = Xc @ K.T
    return q[:, :2] / q[:, 2:3]


def make_problem(n_cams: int, n_points: int, n_obs: int, seed: int = 0, noise_px: float = 0.5,
                 hard: bool = False, windowed: bool = True, arc: int | None = None) -> Problem:
    """Build one synthetic problem.

    windowed=True: each track sees a contiguous window of ring neighbours (video-like);
    windowed=False: uniform-random camera subsets (the survey's stress variant).
    hard=True perturbs the initial guess more (points 0.15, tvec 0.1) so that >= 3 LM iterations run.
    arc: use only the first ``arc`` cameras of the ``n_cams`` ring (same angular spacing, no wrap-around): a bounded
    sample of a large config that keeps its geometry and its observations per camera.
    """
    if n_obs < 2 * n_points:
        raise ValueError("every point needs at least two observations")
    rng = np.random.default_rng(seed)
    K = np.array([[1000.0, 0, 640.0], [0, 1000.0, 360.0], [0, 0, 1.0]])
    ext = ring_cameras(n_cams)
    if arc is not None:
        if not windowed:
            raise ValueError("arc needs windowed visibility")
        ext = ext[:arc]
        n_cams = arc
    X = rng.normal(0.0, 1.0, (n_points, 3))

    base = n_obs // n_points
    lengths = np.full(n_points, base, dtype=np.int64)
    lengths[rng.permutation(n_points)[: n_obs - base * n_points]] += 1
    if lengths.max() > n_cams:
        raise ValueError("track length exceeds the number of cameras")
    pt_idx = np.repeat(np.arange(n_points, dtype=np.int64), lengths)
    offs = np.arange(n_obs, dtype=np.int64) - np.repeat(np.cumsum(lengths) - lengths, lengths)
    if windowed and arc is not None:
        start = rng.integers(n_cams - lengths.max() + 1, size=n_points)
        cam_idx = np.repeat(start, lengths) + offs
    elif windowed:
        start = rng.integers(n_cams, size=n_points)
        cam_idx = (np.repeat(start, lengths) + offs) % n_cams
        # frames inside a track in ascending keyframe order (track.py:12-18)
        order = np.lexsort((cam_idx, pt_idx))
        cam_idx = cam_idx[order]
    else:
        keys = rng.random((n_points, n_cams)) if n_points * n_cams <= 50_000_000 else None
        if keys is not None:
            ranked = np.argsort(keys, axis=1)
            cam_idx = np.sort(np.where(np.arange(n_cams)[None, :] < lengths[:, None], ranked, n_cams),
                              axis=1)
            cam_idx = cam_idx[cam_idx < n_cams].astype(np.int64)
        else:
            cam_idx = np.empty(n_obs, dtype=np.int64)
            pos = 0
            for L in lengths:
                cam_idx[pos:pos + L] = np.sort(rng.choice(n_cams, L, replace=False))
                pos += L

    uv = _project(ext, K, X, cam_idx, pt_idx) + rng.normal(0.0, noise_px, (n_obs, 2))

    sp, st = (0.15, 0.1) if hard else (0.02, 0.01)
    X0 = X + rng.normal(0.0, sp, X.shape)
    ext0 = ext.copy()
    ext0[:, :, 3] += rng.normal(0.0, st, (n_cams, 3))
    return Problem(ext0, K, X0.reshape(n_points, 1, 3), uv, cam_idx, pt_idx, ext, X)


def make_config(name: str, hard: bool = True, windowed: bool = True, scale: float = 1.0, arc: bool = False) -> Problem:
    """One of BASELINE.json's configs; ``scale`` < 1 shrinks points/observations proportionally for bounded CPU
    samples: with ``arc=False`` all cameras are kept (fewer observations per camera), with ``arc=True`` the sample is
    an arc of the camera ring with the config's own observations per camera (the choice for the many-camera configs,
    where keeping every camera would leave most of them with a handful of observations)."""
    nc, npts, nobs = CONFIGS[name]
    if scale != 1.0:
        npts = max(2, int(round(npts * scale)))
        nobs = max(2 * npts, int(round(nobs * scale)))
    if arc and scale != 1.0:
        n_arc = min(nc, max(3 * (nobs // npts + 1), int(round(nc * scale))))
        return make_problem(nc, npts, nobs, seed=CONFIG_SEEDS[name], hard=hard, windowed=True, arc=n_arc)
    return make_problem(nc, npts, nobs, seed=CONFIG_SEEDS[name], hard=hard, windowed=windowed)
