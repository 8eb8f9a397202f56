"""Batched two-view triangulation: oracle vs cv2 golden (CPU), kernel vs both (GPU)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from meatmodeler_b200 import _capi
from meatmodeler_b200 import processor_ops as mp
from oracle import processor_oracle as po       # checker only
from oracle import triangulate_oracle as tri   # checker only


@pytest.fixture(scope="module")
def golden_tri():
    return dict(np.load(os.path.join(GOLDEN, "triangulate.npz")))


class _Track:
    """The reference's track.Track interface (track.py:1-41), re-stated for the tests."""

    def __init__(self, coords, *rest):
        if rest:     # Track(prev_frame_ID, feature, frame_ID, correspondent), track.py:2
            coords = {coords: rest[0], rest[1]: rest[2]}
        self.coordinates = dict(coords)
        self.point = None
        self.updated = False

    def update(self, frame_ID, correspondent):
        self.coordinates[frame_ID] = correspondent
        self.updated = True

    def reset(self):
        self.updated = False

    def wasUpdated(self):
        return self.updated

    def getCoordinate(self, frame_ID):
        return self.coordinates.get(frame_ID)

    def getTriangulationData(self):
        frames = list(self.coordinates.keys())
        return frames[0], frames[-1], self.coordinates.get(frames[0]), self.coordinates.get(frames[-1])

    def getCoordinates(self):
        return self.coordinates

    def setPoint(self, point):
        self.point = point

    def getPoint(self):
        return self.point


def test_oracle_matches_cv2_golden(golden_tri):
    g = golden_tri
    out = tri.triangulate(g["projections"], g["f1"], g["f2"], g["uv1"], g["uv2"])
    assert np.abs(out - g["points"]).max() <= 1e-11 * np.abs(g["points"]).max()


def test_manage_points_matches_reference_order():
    # processor.py:264-291: one observation per (track, frame), track order, dict insertion order
    tracks = [_Track({3: (1.0, 2.0), 5: (3.0, 4.0), 4: (9.0, 9.5)}), _Track({0: (5.0, 6.0), 1: (7.0, 8.0)})]
    tracks[0].setPoint(np.array([[1.0, 2.0, 3.0]]))
    tracks[1].setPoint(np.array([[4.0, 5.0, 6.0]]))
    points, coordinates, frame_indices, point_indices = mp.managePoints(tracks)
    assert (points, coordinates, frame_indices, point_indices) == po.manage_points(tracks)
    assert [p.tolist() for p in points] == [[[1.0, 2.0, 3.0]], [[4.0, 5.0, 6.0]]]
    assert coordinates == [(1.0, 2.0), (3.0, 4.0), (9.0, 9.5), (5.0, 6.0), (7.0, 8.0)]
    assert frame_indices == [3, 5, 4, 0, 1]
    assert point_indices == [0, 0, 0, 1, 1]


def test_point_tracking_matches_the_reference_scan():
    rng = np.random.default_rng(9)

    def scenario():
        tracks = []
        for i in range(300):
            a = (float(rng.integers(0, 60)), float(rng.integers(0, 40)))      # few pixels: duplicates on purpose
            tracks.append(_Track({3: (1.0, 1.0), 4: a} if i % 3 else {2: a, 3: (5.0, 5.0)}))
        feats = np.array([[rng.integers(0, 60), rng.integers(0, 40)] for _ in range(400)], dtype=np.float32)
        feats[7] = feats[3]                                                    # two matches on one feature
        corr = rng.normal(100, 30, (400, 2)).astype(np.float32)
        return tracks, feats, corr

    state = rng.bit_generator.state
    t_ref, feats, corr = scenario()
    rng.bit_generator.state = state
    t_new, feats2, corr2 = scenario()
    assert np.array_equal(feats, feats2)
    popped_ref, upd_ref = po.point_tracking(t_ref, 4, feats, 5, corr, _Track)
    popped_new, upd_new = mp.pointTracking(t_new, 4, feats2, 5, corr2, track_class=_Track)
    assert [t_ref.index(t) for t in popped_ref] == [t_new.index(t) for t in popped_new]
    assert len(upd_ref) == len(upd_new)
    for a, b in zip(upd_ref, upd_new):
        assert a.getCoordinates() == b.getCoordinates() and a.wasUpdated() == b.wasUpdated()
    assert any(len(t.getCoordinates()) == 3 for t in upd_new) and len(popped_new) > 0


def test_install_rebinds_a_processor_module():
    import types
    fake = types.ModuleType("processor")
    fake.Track = _Track
    mp.install(fake)
    assert fake.pointTracking is mp.pointTracking and fake.managePoints is mp.managePoints
    popped, kept = fake.pointTracking([], 0, np.array([[1.0, 2.0]]), 1, np.array([[3.0, 4.0]]))
    assert popped == [] and len(kept) == 1 and kept[0].getCoordinates() == {0: (1.0, 2.0), 1: (3.0, 4.0)}


def test_triangulation_arrays_take_first_and_last_frame():
    t = _Track({7: (1.0, 2.0), 9: (3.0, 4.0), 8: (5.0, 6.0)})
    f1, f2, uv1, uv2 = mp.triangulationArrays([t])
    assert (f1[0], f2[0]) == (7, 8) and uv1[0].tolist() == [1.0, 2.0] and uv2[0].tolist() == [5.0, 6.0]


@pytest.mark.gpu
def test_kernel_vs_cv2_golden_and_oracle(golden_tri):
    g = golden_tri
    out = _capi.triangulate(g["projections"], g["f1"], g["f2"], g["uv1"], g["uv2"])
    scale = np.abs(g["points"]).max()
    assert np.abs(out - g["points"]).max() <= 1e-9 * scale           # bar: 1e-9 relative in float64
    assert np.abs(out - tri.triangulate(g["projections"], g["f1"], g["f2"], g["uv1"], g["uv2"])).max() <= 1e-9 * scale


@pytest.mark.gpu
def test_triangulate_points_dropin(golden_tri):
    g = golden_tri
    n = 50
    tracks = [_Track({int(g["f1"][i]): tuple(g["uv1"][i]), int(g["f2"][i]): tuple(g["uv2"][i])}) for i in range(n)]
    mp.triangulatePoints(tracks, g["projections"])
    for i, t in enumerate(tracks):
        assert t.getPoint().shape == (1, 3)
        assert np.abs(t.getPoint()[0] - g["points"][i]).max() <= 1e-9 * np.abs(g["points"]).max()
    # projections as a {frame_ID: P} mapping
    tracks2 = [_Track({int(g["f1"][i]): tuple(g["uv1"][i]), int(g["f2"][i]): tuple(g["uv2"][i])}) for i in range(n)]
    mp.triangulatePoints(tracks2, {k: P for k, P in enumerate(g["projections"])})
    assert all(np.array_equal(a.getPoint(), b.getPoint()) for a, b in zip(tracks, tracks2))
    mp.triangulatePoints([], g["projections"])   # empty: no-op, as the reference loop


@pytest.mark.gpu
def test_triangulate_edge_cases_and_large_batch(golden_tri):
    g = golden_tri
    assert _capi.triangulate(g["projections"], [], [], np.zeros((0, 2)), np.zeros((0, 2))).shape == (0, 3)
    with pytest.raises(_capi.MmbaError):
        _capi.triangulate(g["projections"], [len(g["projections"])], [0], np.zeros((1, 2)), np.zeros((1, 2)))
    # 1M tracks (noise-free, two exact views): the triangulated point re-projects onto both pixels
    rng = np.random.default_rng(5)
    P = g["projections"]
    n = 1_000_000
    X = rng.normal(0, 1, (n, 3))
    f1 = rng.integers(0, len(P), n)
    f2 = (f1 + rng.integers(1, 6, n)) % len(P)
    Xh = np.c_[X, np.ones(n)]
    q1 = np.einsum("nij,nj->ni", P[f1], Xh)
    q2 = np.einsum("nij,nj->ni", P[f2], Xh)
    uv1, uv2 = q1[:, :2] / q1[:, 2:], q2[:, :2] / q2[:, 2:]
    out, ms = _capi.triangulate(P, f1, f2, uv1, uv2, return_ms=True)
    assert np.abs(out - X).max() <= 1e-7
    assert ms > 0


# ---- pinned against the UNMODIFIED reference (tests/golden/make_golden_processor.py -> processor.npz) -----------------

@pytest.fixture(scope="module")
def golden_proc():
    return dict(np.load(os.path.join(GOLDEN, "processor.npz")))


def _unpack_tracks(ptr, frames, xy, as_f32=True):
    out = []
    for i in range(len(ptr) - 1):
        coords = {}
        for j in range(ptr[i], ptr[i + 1]):
            c = (np.float32(xy[j, 0]), np.float32(xy[j, 1])) if as_f32 else (float(xy[j, 0]), float(xy[j, 1]))
            coords[int(frames[j])] = c
        out.append(_Track(coords))
    return out


def _pack(tracks):
    ptr, frames, xy = [0], [], []
    for t in tracks:
        for f, c in t.getCoordinates().items():
            frames.append(f)
            xy.append((float(c[0]), float(c[1])))
        ptr.append(len(frames))
    return np.array(ptr), np.array(frames), np.array(xy, dtype=np.float64).reshape(-1, 2)


@pytest.mark.parametrize("impl", ["oracle", "product"])
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_point_tracking_vs_reference_golden(golden_proc, tag, impl):
    """pointTracking of the unmodified reference (processor.py:190-243) on seeded scenarios: which tracks are popped, which
    survive (in order), what the survivors and the new tracks contain — for the oracle restatement AND the hash join."""
    g = golden_proc
    tracks = _unpack_tracks(g[f"pt_{tag}_in_ptr"], g[f"pt_{tag}_in_frames"], g[f"pt_{tag}_in_xy"])
    feats, corr = g[f"pt_{tag}_feats"], g[f"pt_{tag}_corr"]
    if impl == "oracle":
        popped, updated = po.point_tracking(tracks, 4, feats, 5, corr, _Track)
    else:
        popped, updated = mp.pointTracking(tracks, 4, feats, 5, corr, track_class=_Track)
    ids = {id(t): i for i, t in enumerate(tracks)}
    np.testing.assert_array_equal([ids[id(t)] for t in popped], g[f"pt_{tag}_popped"])
    np.testing.assert_array_equal([ids.get(id(t), -1) for t in updated], g[f"pt_{tag}_updated_src"])
    ptr, frames, xy = _pack(updated)
    np.testing.assert_array_equal(ptr, g[f"pt_{tag}_up_ptr"])
    np.testing.assert_array_equal(frames, g[f"pt_{tag}_up_frames"])
    np.testing.assert_array_equal(xy, g[f"pt_{tag}_up_xy"])
    np.testing.assert_array_equal([t.wasUpdated() for t in updated], g[f"pt_{tag}_updated_flags"])


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_manage_points_vs_reference_golden(golden_proc, impl):
    g = golden_proc
    tracks = _unpack_tracks(g["mp_ptr"], g["mp_frames"], g["mp_xy"], as_f32=False)
    for t, p in zip(tracks, g["mp_points"]):
        t.setPoint(p.reshape(tuple(g["mp_point_shape"])))
    fn = po.manage_points if impl == "oracle" else mp.managePoints
    points, coordinates, frame_indices, point_indices = fn(tracks)
    np.testing.assert_array_equal(np.array(points).reshape(-1, 3), g["mp_points"])
    assert np.array(points[0]).shape == tuple(g["mp_point_shape"])
    np.testing.assert_array_equal(np.array(coordinates, dtype=np.float64), g["mp_coordinates"])
    np.testing.assert_array_equal(frame_indices, g["mp_frame_indices"])
    np.testing.assert_array_equal(point_indices, g["mp_point_indices"])


def test_triangulate_oracle_vs_reference_processor_golden(golden_proc):
    """The oracle's DLT against the points the unmodified processor.triangulatePoints stored on its tracks."""
    g = golden_proc
    ptr, frames, xy = g["mp_ptr"], g["mp_frames"], g["mp_xy"]
    f1, f2 = frames[ptr[:-1]], frames[ptr[1:] - 1]
    out = tri.triangulate(g["mp_projections"], f1, f2, xy[ptr[:-1]], xy[ptr[1:] - 1])
    assert np.abs(out - g["mp_points"]).max() <= 1e-9 * np.abs(g["mp_points"]).max()


@pytest.mark.gpu
def test_triangulate_points_dropin_vs_reference_processor_golden(golden_proc):
    """processor_ops.triangulatePoints (one kernel launch over all tracks) against the unmodified reference's per-track
    cv2 loop: same points on the tracks, same (1, 3) shape."""
    g = golden_proc
    tracks = _unpack_tracks(g["mp_ptr"], g["mp_frames"], g["mp_xy"], as_f32=False)
    mp.triangulatePoints(tracks, list(g["mp_projections"]))
    got = np.array([t.getPoint() for t in tracks])
    assert got.shape[1:] == tuple(g["mp_point_shape"])
    assert np.abs(got.reshape(-1, 3) - g["mp_points"]).max() <= 1e-9 * np.abs(g["mp_points"]).max()


def test_save_point_cloud_writes_a_binary_ply(tmp_path):
    """processor.py:480-485 (PyntCloud(...).to_file): header + raw float64 records; read back with a minimal parser."""
    pts = np.random.default_rng(3).normal(size=(257, 3))
    path = mp.savePointCloud(pts, str(tmp_path / "Cloud.ply"))
    raw = open(path, "rb").read()
    head, _, body = raw.partition(b"end_header\n")
    lines = head.decode().splitlines()
    assert lines[0] == "ply" and lines[1].startswith("format binary_") and lines[1].endswith("_endian 1.0")
    assert lines[2] == "element vertex 257" and lines[3:6] == ["property double x", "property double y", "property double z"]
    np.testing.assert_array_equal(np.frombuffer(body, dtype=np.float64).reshape(-1, 3), pts)
