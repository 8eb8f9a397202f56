"""Golden vectors of the UNMODIFIED reference ``processor.py`` steps next to the hot path:
``pointTracking`` (processor.py:190-243), ``managePoints`` (:264-291) and ``triangulatePoints`` (:246-261), run on seeded
scenarios with the reference's own ``Track`` class (track.py).

Build container only (needs /root/reference).  ``processor.py`` imports pyntcloud / lxml for its PLY export; neither is
installed in this image and neither is touched by the three functions, so empty stand-in modules are registered
before the import.  Writes tests/golden/processor.npz; nothing at test time reads /root/reference.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for name in ("pyntcloud", "lxml"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["pyntcloud"].PyntCloud = object
sys.path.insert(0, "/root/reference")
import processor as ref          # noqa: E402  (the reference, unmodified)
from track import Track          # noqa: E402


def pack_tracks(tracks):
    """tracks -> (ptr, frames, xy): coordinate dictionaries in insertion order."""
    ptr, frames, xy = [0], [], []
    for t in tracks:
        for f, c in t.getCoordinates().items():
            frames.append(f)
            xy.append((float(c[0]), float(c[1])))
        ptr.append(len(frames))
    return np.array(ptr, dtype=np.int64), np.array(frames, dtype=np.int64), np.array(xy, dtype=np.float64).reshape(-1, 2)


def tracking_scenario(seed, n_tracks, n_matches, prev_id, cur_id):
    rng = np.random.default_rng(seed)
    tracks = []
    for i in range(n_tracks):
        a = (np.float32(rng.integers(0, 60)), np.float32(rng.integers(0, 40)))      # few pixels: duplicates on purpose
        if i % 3:
            t = Track(prev_id - 1, (np.float32(1.0), np.float32(1.0)), prev_id, a)
        else:
            t = Track(prev_id - 2, a, prev_id - 1, (np.float32(5.0), np.float32(5.0)))   # not seen in the previous keyframe
        tracks.append(t)
    feats = np.array([[rng.integers(0, 60), rng.integers(0, 40)] for _ in range(n_matches)], dtype=np.float32)
    feats[7] = feats[3]                                                                # two matches on one feature
    corr = rng.normal(100, 30, (n_matches, 2)).astype(np.float32)
    return tracks, feats, corr


def main():
    out = {}
    # ---- pointTracking -----------------------------------------------------------------------
    for tag, (seed, nt, nm) in {"a": (9, 300, 400), "b": (10, 50, 20), "c": (11, 5, 64)}.items():
        tracks, feats, corr = tracking_scenario(seed, nt, nm, 4, 5)
        in_ptr, in_frames, in_xy = pack_tracks(tracks)
        ids = {id(t): i for i, t in enumerate(tracks)}
        popped, updated = ref.pointTracking(tracks, 4, feats, 5, corr)
        up_ptr, up_frames, up_xy = pack_tracks(updated)
        out.update({f"pt_{tag}_in_ptr": in_ptr, f"pt_{tag}_in_frames": in_frames, f"pt_{tag}_in_xy": in_xy,
                    f"pt_{tag}_feats": feats, f"pt_{tag}_corr": corr,
                    f"pt_{tag}_popped": np.array([ids[id(t)] for t in popped], dtype=np.int64),
                    f"pt_{tag}_updated_src": np.array([ids.get(id(t), -1) for t in updated], dtype=np.int64),
                    f"pt_{tag}_up_ptr": up_ptr, f"pt_{tag}_up_frames": up_frames, f"pt_{tag}_up_xy": up_xy,
                    f"pt_{tag}_updated_flags": np.array([t.wasUpdated() for t in updated])})
    # ---- triangulatePoints + managePoints on the same tracks ----------------------------------
    rng = np.random.default_rng(21)
    n_frames, n_tracks = 12, 200
    K = np.array([[1000.0, 0, 640.0], [0, 1000.0, 360.0], [0, 0, 1.0]])
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from meatmodeler_b200 import synth
    ext = synth.ring_cameras(n_frames)
    projections = [K @ ext[i] for i in range(n_frames)]
    X = rng.normal(0, 1, (n_tracks, 3))
    tracks = []
    for p in range(n_tracks):
        L = int(rng.integers(2, 7))
        start = int(rng.integers(0, n_frames - L + 1))
        t = None
        for f in range(start, start + L):
            q = projections[f] @ np.append(X[p], 1.0)
            # float32 pixel values (what the optical flow returns) as Python floats: cv2 4.13 rejects tuples of numpy scalars
            c = (float(np.float32(q[0] / q[2] + rng.normal(0, 0.3))), float(np.float32(q[1] / q[2] + rng.normal(0, 0.3))))
            if t is None:
                first = (f, c)
                t = False
            elif t is False:
                t = Track(first[0], first[1], f, c)
            else:
                t.update(f, c)
        tracks.append(t)
    ref.triangulatePoints(tracks, projections)
    points, coordinates, frame_indices, point_indices = ref.managePoints(tracks)
    ptr, frames, xy = pack_tracks(tracks)
    out.update(mp_projections=np.array(projections), mp_ptr=ptr, mp_frames=frames, mp_xy=xy,
               mp_points=np.array(points).reshape(-1, 3), mp_point_shape=np.array(np.array(points[0]).shape),
               mp_coordinates=np.array(coordinates, dtype=np.float64), mp_frame_indices=np.array(frame_indices, dtype=np.int64),
               mp_point_indices=np.array(point_indices, dtype=np.int64))
    np.savez_compressed(os.path.join(HERE, "processor.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
