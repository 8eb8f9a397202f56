"""Drop-in for the two ``processor.py`` steps either side of the bundle-adjustment hot path:

* ``triangulatePoints(tracks, projections)``  (processor.py:246-261) — the per-track
  ``cv2.triangulatePoints`` loop becomes ONE launch of ``triangulate_kernel`` (libmmba.so) over all tracks;
* ``managePoints(tracks)``                    (processor.py:264-291) — the observation lists
  (points, coordinates, frame indices, point indices) that feed ``bundleAdjuster.adjustPoints``.

Same names, argument meaning and side effects as the reference (``track.setPoint`` receives a (1, 3) array, as
processor.py:259-260 produces).  ``tracks`` are duck-typed: any object with the methods of the reference's
``track.Track`` (track.py:1-41).  No CPU fallback: the triangulation fails loudly without libmmba.so / a B200.
"""
import numpy as np

from . import _capi


def triangulationArrays(tracks):
    """First / last frame and pixel of every track (track.getTriangulationData, track.py:31-33) as flat arrays."""
    n = len(tracks)
    f1 = np.empty(n, dtype=np.int64)
    f2 = np.empty(n, dtype=np.int64)
    uv1 = np.empty((n, 2), dtype=np.float64)
    uv2 = np.empty((n, 2), dtype=np.float64)
    for i, track in enumerate(tracks):
        a, b, feature, correspondent = track.getTriangulationData()
        f1[i], f2[i] = a, b
        uv1[i] = np.asarray(feature, dtype=np.float64).reshape(-1)[:2]
        uv2[i] = np.asarray(correspondent, dtype=np.float64).reshape(-1)[:2]
    return f1, f2, uv1, uv2


def triangulatePoints(tracks, projections, device=0):
    """processor.py:246-261.  ``projections``: sequence (or dict) of 3x4 projection matrices indexed by frame ID."""
    tracks = list(tracks)
    if not tracks:
        return
    f1, f2, uv1, uv2 = triangulationArrays(tracks)
    if isinstance(projections, dict):
        ids = np.unique(np.concatenate((f1, f2)))
        proj = np.stack([np.asarray(projections[int(k)], dtype=np.float64) for k in ids])
        f1 = np.searchsorted(ids, f1)
        f2 = np.searchsorted(ids, f2)
    else:
        proj = np.asarray(projections, dtype=np.float64)
    points = _capi.triangulate(proj, f1, f2, uv1, uv2, device=device)
    for track, point in zip(tracks, points):
        track.setPoint(point.reshape(1, 3))


def managePoints(tracks):
    """processor.py:264-291: (points, coordinates, frame_indices, point_indices) in track order, one observation
    per (track, frame) in the insertion order of ``track.getCoordinates()``."""
    points = []
    coordinates = []
    point_indices = []
    frame_indices = []
    for point_index, track in enumerate(tracks):
        points.append(track.getPoint())
        coords = track.getCoordinates()
        coordinates.extend(coords.values())
        frame_indices.extend(coords.keys())
        point_indices.extend([point_index] * len(coords))
    return points, coordinates, frame_indices, point_indices
