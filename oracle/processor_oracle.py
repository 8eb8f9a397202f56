"""CPU restatement of the ``processor.py`` steps next to the bundle-adjustment path.  TEST INFRASTRUCTURE
ONLY: nothing under meatmodeler_b200/ imports this module.

Parity status: PINNED.  tests/golden/make_golden_processor.py imports the unmodified ``processor.py`` (its pyntcloud /
lxml imports, needed only by the PLY export, are satisfied by empty stand-in modules) and records what ``pointTracking``,
``triangulatePoints`` and ``managePoints`` do on seeded scenarios with the reference's own ``Track`` class;
tests/test_processor_ops.py checks this restatement and the product (meatmodeler_b200/processor_ops.py) against those
vectors (tests/golden/processor.npz).
"""


def point_tracking(tracks, prev_keyframe_ID, feature_points, keyframe_ID, correspondents, make_track):
    """processor.py:190-243: for every match scan ALL tracks for the first one whose pixel in the previous keyframe
    equals the feature point (tuple equality, :218) and update it (:219-221), else start a new track (:224-229);
    afterwards updated tracks are reset and kept, the others popped (:231-241)."""
    new_tracks, updated_tracks, popped_tracks = [], [], []
    for feature_point, correspondent in zip(feature_points, correspondents):
        feature_point = (feature_point[0], feature_point[1])
        correspondent = (correspondent[0], correspondent[1])
        is_new_track = True
        for track in tracks:
            if feature_point == track.getCoordinate(prev_keyframe_ID):
                track.update(keyframe_ID, correspondent)
                is_new_track = False
                break
        if is_new_track:
            new_tracks.append(make_track(prev_keyframe_ID, feature_point, keyframe_ID, correspondent))
    for track in tracks:
        if track.wasUpdated():
            track.reset()
            updated_tracks.append(track)
        else:
            popped_tracks.append(track)
    return popped_tracks, updated_tracks + new_tracks


def manage_points(tracks):
    """processor.py:264-291: points in track order; one (coordinate, frame index, point index) per (track, frame) in
    the insertion order of the track's coordinate dictionary."""
    points, coordinates, frame_indices, point_indices = [], [], [], []
    for point_index, track in enumerate(tracks):
        points.append(track.getPoint())
        for frame_index, coordinate in track.getCoordinates().items():
            coordinates.append(coordinate)
            point_indices.append(point_index)
            frame_indices.append(frame_index)
    return points, coordinates, frame_indices, point_indices
