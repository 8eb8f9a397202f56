"""CPU oracle for the batched two-view triangulation.  TEST INFRASTRUCTURE ONLY: nothing
under meatmodeler_b200/ imports this module.

Restates what the reference does per track (processor.py:246-261): ``cv2.triangulatePoints(P1, P2, x1, x2)``
followed by de-homogenisation (processor.py:259).  OpenCV (pinned ~=4.5.2 by the reference's requirements.txt,
4.13 in the build container; the algorithm lives in third-party modules/calib3d/src/triangulate.cpp, not
under /root/reference) builds the 4x4 DLT matrix
    A = [x1*P1[2] - P1[0]; y1*P1[2] - P1[1]; x2*P2[2] - P2[0]; y2*P2[2] - P2[1]]
and returns the right singular vector of its smallest singular value.

Pinning: tests/golden/triangulate.npz holds outputs of cv2.triangulatePoints itself called exactly as the
reference loop calls it (tests/golden/make_golden_tri.py); tests/test_oracle_golden.py checks this file
against them.
"""
import numpy as np


def dlt_matrix(P1, P2, x1, x2):
    """processor.py:255-258 -> the 4x4 system cv2.triangulatePoints assembles for one track."""
    P1 = np.asarray(P1, dtype=np.float64)
    P2 = np.asarray(P2, dtype=np.float64)
    return np.stack((x1[0] * P1[2] - P1[0], x1[1] * P1[2] - P1[1], x2[0] * P2[2] - P2[0], x2[1] * P2[2] - P2[1]))


def triangulate(projections, f1, f2, uv1, uv2):
    """(n,3) points: per-track DLT + de-homogenisation (processor.py:255-259), one SVD per track."""
    projections = np.asarray(projections, dtype=np.float64).reshape(-1, 3, 4)
    uv1 = np.asarray(uv1, dtype=np.float64).reshape(-1, 2)
    uv2 = np.asarray(uv2, dtype=np.float64).reshape(-1, 2)
    out = np.empty((len(uv1), 3))
    for i in range(len(uv1)):
        A = dlt_matrix(projections[f1[i]], projections[f2[i]], uv1[i], uv2[i])
        X = np.linalg.svd(A)[2][-1]
        out[i] = X[:3] / X[3]
    return out
