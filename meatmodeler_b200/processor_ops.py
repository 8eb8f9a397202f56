"""Drop-in for the ``processor.py`` steps either side of the bundle-adjustment hot path:

* ``pointTracking(tracks, prev_keyframe_ID, feature_points, keyframe_ID, correspondents)`` (processor.py:190-243) —
  the O(matches x tracks) scan becomes one hash join (first track per previous-keyframe pixel);
* ``triangulatePoints(tracks, projections)``  (processor.py:246-261) — the per-track
  ``cv2.triangulatePoints`` loop becomes ONE launch of ``triangulate_kernel`` (libmmba.so) over all tracks;
* ``managePoints(tracks)``                    (processor.py:264-291) — the observation lists
  (points, coordinates, frame indices, point indices) that feed ``bundleAdjuster.adjustPoints``.

Same names, argument meaning, return values and side effects as the reference (``track.setPoint`` receives a (1, 3)
array, as processor.py:259-260 produces).  ``tracks`` are duck-typed: any object with the methods of the reference's
``track.Track`` (track.py:1-41); new tracks are built with the caller's own ``Track`` class (``install`` takes it from
the patched module).  No CPU fallback: the triangulation fails loudly without libmmba.so / a B200.

    import processor, meatmodeler_b200.processor as mp
    mp.install(processor)        # processor.pointTracking / triangulatePoints / managePoints now resolve here
"""
import numpy as np

from . import _capi

_track_class = None


def install(module, track_class=None):
    """Rebind the three functions of an imported ``processor`` module (the reference's, unmodified) to this module."""
    global _track_class
    _track_class = track_class if track_class is not None else getattr(module, "Track", None)
    module.pointTracking = pointTracking
    module.triangulatePoints = triangulatePoints
    module.managePoints = managePoints
    return module


def pointTracking(tracks, prev_keyframe_ID, feature_points, keyframe_ID, correspondents, track_class=None):
    """processor.py:190-243.  A match continues the FIRST track (list order) whose pixel in the previous keyframe
    equals the match's feature point, otherwise it starts a new track; tracks that received no match are popped.
    Returns (popped_tracks, updated_tracks + new_tracks), both in the reference's order."""
    make = track_class or _track_class
    if make is None:
        raise TypeError("pointTracking needs the Track class: pass track_class= or call install(processor) first")
    first = {}
    for track in tracks:
        prior = track.getCoordinate(prev_keyframe_ID)
        if prior is not None:
            first.setdefault(_pixel_key(prior), track)
    new_tracks = []
    for feature_point, correspondent in zip(feature_points, correspondents):
        feature_point = (feature_point[0], feature_point[1])
        correspondent = (correspondent[0], correspondent[1])
        track = first.get(_pixel_key(feature_point))
        if track is not None:
            track.update(keyframe_ID, correspondent)
        else:
            new_tracks.append(make(prev_keyframe_ID, feature_point, keyframe_ID, correspondent))
    updated_tracks = []
    popped_tracks = []
    for track in tracks:
        if track.wasUpdated():
            track.reset()
            updated_tracks.append(track)
        else:
            popped_tracks.append(track)
    updated_tracks += new_tracks
    return popped_tracks, updated_tracks


def _pixel_key(p):
    # tuple equality of the reference (processor.py:218): == on both coordinates; -0.0 == 0.0, nan never matches
    x, y = float(p[0]), float(p[1])
    if x != x or y != y:
        return object()
    return (x + 0.0, y + 0.0)


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


def savePointCloud(points, filename):
    """The output side of ``processor.process`` (processor.py:480-485):
    ``PyntCloud(pd.DataFrame(adjusted_points, columns=['x', 'y', 'z'])).to_file(filename)`` without pandas / pyntcloud.

    Writes the file pyntcloud's PLY writer produces for that frame (pyntcloud/io/ply.py ``write_ply``, ``as_text=False``):
    a text header — ``ply``, ``format binary_<byteorder>_endian 1.0``, ``element vertex N``, one ``property double``
    line per column, ``end_header`` — followed by the N x 3 float64 records, i.e. the result buffer of ``adjustPoints``
    as it is.  pyntcloud is not installed in this image, so the format is restated from its source, not pinned.
    """
    import sys
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    header = ["ply", f"format binary_{sys.byteorder}_endian 1.0", f"element vertex {len(pts)}",
              "property double x", "property double y", "property double z", "end_header"]
    with open(filename, "w") as f:
        for line in header:
            f.write(f"{line}\n")
    with open(filename, "ab") as f:
        pts.tofile(f)
    return filename
