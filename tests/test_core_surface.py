"""The reference-shaped Python surface (`polychase_core`, polychase_pybind.cc:29-348): names,
defaults and the SQLite format.  CPU only (no kernels are launched)."""
import os

import numpy as np
import pytest

from oracle import db as odb


@pytest.fixture(scope="module")
def core():
    from polychase_b200 import build_pybind
    build_pybind.build()
    from polychase_b200 import polychase_core
    return polychase_core


def test_surface_names(core):
    for name in ("Mesh", "AcceleratedMesh", "SceneTransformations", "RayHit", "Database", "ImagePairFlow", "VideoInfo",
                 "GFTTOptions", "OpticalFlowOptions", "TrackerThread", "RefinerThread", "OpticalFlowProgress",
                 "OpticalFlowRequest", "OpticalFlowThread", "TransformationType", "CameraConvention",
                 "CameraIntrinsics", "Pose", "CameraState", "LossType", "BundleOptions", "BundleStats", "PnPResult",
                 "FrameTrackingResult", "CameraTrajectory", "RefineTrajectoryUpdate", "CppException", "ray_cast",
                 "generate_optical_flow_database", "track_sequence", "refine_trajectory"):
        assert hasattr(core, name), name


def test_defaults_match_reference(core):
    g = core.GFTTOptions()                       # gftt.h:5-21
    assert (g.quality_level, g.min_distance, g.block_size, g.gradient_size, g.max_corners, g.use_harris,
            g.harris_k) == (0.01, 5.0, 3, 3, 0, False, 0.04)
    assert not hasattr(g, "grid_rows")           # not exposed (polychase_pybind.cc:128-136)
    f = core.OpticalFlowOptions()                # opticalflow.h:27-33
    assert (f.window_size, f.max_level, f.term_max_iters, f.term_epsilon, f.min_eigen_threshold) == (10, 3, 30, 0.01, 1e-4)
    b = core.BundleOptions()                     # types.h:200-215
    assert b.max_iterations == 100 and b.max_allowed_parallelism == 8 and b.loss_type == core.LossType.Huber
    assert abs(b.initial_lambda - 1e-5) < 1e-12 and abs(b.max_lambda - 1e10) < 1
    it = core.CameraIntrinsics(fx=1, fy=2, cx=3, cy=4, aspect_ratio=1, width=10, height=20)
    assert it.convention == core.CameraConvention.OpenGL
    p = core.Pose()
    assert list(p.q) == [1, 0, 0, 0] and list(p.t) == [0, 0, 0]        # identity, q as WXYZ
    p.q = np.array([0, 1, 0, 0], np.float32)
    assert list(p.q) == [0, 1, 0, 0]


def test_camera_trajectory(core):
    t = core.CameraTrajectory(first_frame_id=5, count=3)
    assert (t.count(), t.first_frame(), t.last_frame()) == (3, 5, 7)
    assert t.is_valid_frame(7) and not t.is_valid_frame(8) and not t.is_valid_frame(4)
    assert not t.is_frame_filled(6) and t.get(6) is None
    cs = core.CameraState(core.CameraIntrinsics(1, 1, 0, 0, 1, 4, 4), core.Pose())
    t.set(6, cs)
    assert t.is_frame_filled(6) and t.get(6).intrinsics.width == 4
    with pytest.raises(Exception):               # CHECK(index < Count()) -> std::logic_error
        t.get(9)


def test_mesh_mask_bits(core):
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], np.float32)
    t = np.array([[0, 1, 2], [1, 3, 2]], np.uint32)
    m = core.AcceleratedMesh(v, t)
    inner = m.inner_mut()
    assert len(inner.masked_triangles) == 4      # padded to a multiple of 4 words (geometry.h:63-65)
    assert not inner.is_triangle_masked(1)
    inner.mask_triangle(1)
    assert inner.is_triangle_masked(1)
    inner.toggle_mask_triangle(1)
    assert not m.inner().is_triangle_masked(1)
    with pytest.raises(Exception):
        core.AcceleratedMesh(v, np.array([[0, 1, 9]], np.uint32))


def test_database_format_is_the_reference_format(core, tmp_path):
    """Rows written by the new Database read back identically through a plain sqlite3 restatement of
    database.cc, and vice versa (blob-identical)."""
    rng = np.random.default_rng(0)
    kps = rng.uniform(0, 100, (7, 2)).astype(np.float32)
    idx = np.array([0, 2, 5], np.uint32)
    tgt = rng.uniform(0, 100, (3, 2)).astype(np.float32)
    err = rng.uniform(0, 1, 3).astype(np.float32)
    p1 = str(tmp_path / "a.db")
    d = core.Database(p1)
    d.write_keypoints(3, kps)
    d.write_keypoints(4, kps[:2])
    d.write_image_pair_flow(3, 4, idx, tgt, err)
    assert d.keypoints_exist(3) and not d.keypoints_exist(9)
    assert d.image_pair_flow_exists(3, 4) and not d.image_pair_flow_exists(4, 3)
    assert d.get_min_image_id_with_keypoints() == 3 and d.get_max_image_id_with_keypoints() == 4
    assert d.find_optical_flows_from_image(3) == [4] and d.find_optical_flows_to_image(4) == [3]
    assert np.array_equal(d.read_keypoints(3), kps)
    assert d.read_keypoints(77).shape == (0, 2)
    f = d.read_image_pair_flow(3, 4)
    assert (f.image_id_from, f.image_id_to) == (3, 4)
    assert np.array_equal(f.src_kps_indices, idx) and np.array_equal(f.tgt_kps, tgt) and np.array_equal(f.flow_errors, err)
    with pytest.raises(Exception):               # duplicate primary key -> SQLite error -> runtime_error
        d.write_keypoints(3, kps)
    d.close()
    o = odb.Database(p1)
    assert np.array_equal(o.read_keypoints(3), kps)
    oi, ot, oe = o.read_image_pair_flow(3, 4)
    assert np.array_equal(oi, idx) and np.array_equal(ot, tgt) and np.array_equal(oe, err)
    o.close()
    # and the other direction
    p2 = str(tmp_path / "b.db")
    o = odb.Database(p2)
    o.write_keypoints(1, kps)
    o.write_image_pair_flow(1, 1 + 8, idx, tgt, err) if False else None
    o.write_keypoints(9, kps)
    o.write_image_pair_flow(1, 9, idx, tgt, err)
    o.close()
    d = core.Database(p2)
    assert np.array_equal(d.read_keypoints(1), kps)
    f = d.read_image_pair_flow(1, 9)
    assert np.array_equal(f.src_kps_indices, idx) and np.array_equal(f.tgt_kps, tgt)
    d.close()
    # raw bytes of the two files' blobs agree
    import sqlite3
    a = sqlite3.connect(p1).execute("SELECT keypoints FROM keypoints WHERE image_id=3").fetchone()[0]
    b = sqlite3.connect(p2).execute("SELECT keypoints FROM keypoints WHERE image_id=1").fetchone()[0]
    assert bytes(a) == bytes(b) == kps.tobytes()


def test_empty_blobs_round_trip(core, tmp_path):
    d = core.Database(str(tmp_path / "e.db"))
    d.write_keypoints(1, np.zeros((0, 2), np.float32))
    d.write_image_pair_flow(1, 2, np.zeros(0, np.uint32), np.zeros((0, 2), np.float32), np.zeros(0, np.float32))
    assert d.read_keypoints(1).shape == (0, 2)
    assert len(d.read_image_pair_flow(1, 2).src_kps_indices) == 0
