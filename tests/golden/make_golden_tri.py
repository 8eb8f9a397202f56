"""Golden vectors for the batched triangulation from cv2.triangulatePoints itself.

Run in the build container only (needs cv2):  python tests/golden/make_golden_tri.py
``processor.py`` cannot be imported here (pyntcloud is not installed), so the per-track loop of
``processor.triangulatePoints`` (processor.py:254-260) is replayed verbatim on synthetic tracks:
    point = cv2.triangulatePoints(projection1, projection2, feature, correspondent).T
    point = point[:, :3] / point[:, -1, np.newaxis]
Writes tests/golden/triangulate.npz (projections, first/last frame of each track, the two pixel
coordinates, the reference points, cv2 version).
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from meatmodeler_b200 import synth  # noqa: E402


def main():
    rng = np.random.default_rng(77)
    prob = synth.make_problem(24, 600, 3600, seed=7, hard=False)
    ext, K, pts, uv, fi, pi = prob.args()
    projections = np.stack([K @ e[:3, :4] for e in ext])          # processor.py:183
    order = np.lexsort((fi, pi))
    fi, pi, uv = fi[order], pi[order], uv[order]
    starts = np.flatnonzero(np.r_[True, pi[1:] != pi[:-1]])
    ends = np.r_[starts[1:], len(pi)] - 1
    f1, f2 = fi[starts], fi[ends]
    uv1, uv2 = uv[starts], uv[ends]
    # a few degenerate / hard tracks: tiny baseline (neighbouring frames), noise-free, far point
    out = np.empty((len(f1), 3))
    for i in range(len(f1)):
        feature = (uv1[i, 0], uv1[i, 1])                         # tuples, as pointTracking stores them
        correspondent = (uv2[i, 0], uv2[i, 1])
        point = cv2.triangulatePoints(projections[f1[i]], projections[f2[i]], feature, correspondent).T
        point = point[:, :3] / point[:, -1, np.newaxis]
        out[i] = point[0]
    np.savez_compressed(os.path.join(HERE, "triangulate.npz"), projections=projections, f1=f1, f2=f2, uv1=uv1, uv2=uv2,
                        points=out, truth=np.asarray(pts).reshape(-1, 3)[pi[starts]], cv2_version=cv2.__version__)
    print("tracks", len(f1), "cv2", cv2.__version__, "max |X - truth|", np.abs(out - np.asarray(pts).reshape(-1, 3)[pi[starts]]).max())
    del rng


if __name__ == "__main__":
    main()
