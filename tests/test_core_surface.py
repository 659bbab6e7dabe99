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
                 "generate_optical_flow_database", "track_sequence", "refine_trajectory", "PinUpdate",
                 "find_transformation"):
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


def _scene(core, conv_gl=True):
    w, h, f = 640.0, 480.0, 700.0
    if conv_gl:      # the addon's camera: OpenGL, negated fx / fy (blender_addon/core.py:348-357)
        intr = core.CameraIntrinsics(-f, -f, w / 2, h / 2, 1.0, w, h, core.CameraConvention.OpenGL)
        view = np.diag([1.0, -1.0, -1.0, 1.0]).astype(np.float32)
    else:
        intr = core.CameraIntrinsics(f, f, w / 2, h / 2, 1.0, w, h, core.CameraConvention.OpenCV)
        view = np.eye(4, dtype=np.float32)
    view[:3, 3] = view[:3, :3] @ np.array([0.1, -0.2, 5.0], np.float32)
    model = np.eye(4, dtype=np.float32)
    model[:3, 3] = [0.3, 0.1, 0.2]
    return core.SceneTransformations(model, view, intr)


def _project(scene, pts):
    mv = np.asarray(scene.view_matrix, np.float64) @ np.asarray(scene.model_matrix, np.float64)
    pc = pts @ mv[:3, :3].T + mv[:3, 3]
    k = scene.intrinsics
    return np.stack([k.fx * pc[:, 0] / pc[:, 2] + k.cx, k.fy * pc[:, 1] / pc[:, 2] + k.cy], 1)


def _two_pin_restatement(scene, pts, pin, pos, trans):
    """FindTransformation2 (pin_mode.cc:151-217) written independently in float64 numpy."""
    M, V, k = np.asarray(scene.model_matrix, np.float64), np.asarray(scene.view_matrix, np.float64), scene.intrinsics
    Vi = np.linalg.inv(V)
    s = 1.0 if k.convention.name == "OpenCV" else -1.0
    d = Vi[:3, :3] @ (s * np.array([(pos[0] - k.cx) / k.fx, (pos[1] - k.cy) / k.fy, 1.0]))
    o = Vi[:3, 3]
    moving = M[:3, :3] @ pts[pin] + M[:3, 3]
    anchor = M[:3, :3] @ pts[1 - pin] + M[:3, 3]
    depth = np.linalg.norm(moving - o)
    tm = o + depth * d / np.linalg.norm(d)
    du, dv = moving - anchor, tm - anchor
    dn = Vi[:3, 2] / np.linalg.norm(Vi[:3, 2])
    duu, dvu = du / np.linalg.norm(du), dv / np.linalg.norm(dv)
    ang = np.arctan2(np.cross(duu, dvu) @ dn, duu @ dvu)
    Kx = np.array([[0, -dn[2], dn[1]], [dn[2], 0, -dn[0]], [-dn[1], dn[0], 0]])
    R = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx
    new_anchor = o + (anchor - o) * (np.linalg.norm(du) / np.linalg.norm(dv))
    U = np.eye(4)
    U[:3, :3] = R
    U[:3, 3] = new_anchor - R @ anchor
    return (U @ M, V) if trans == "Model" else (M, V @ U)


@pytest.mark.parametrize("conv_gl", [True, False])
@pytest.mark.parametrize("trans", ["Model", "Camera"])
def test_find_transformation_one_and_two_pins(core, conv_gl, trans):
    """FindTransformation1 / FindTransformation2 (pin_mode.cc:110-217) are closed-form host math: after the update the
    dragged pin projects onto the cursor; with two pins the anchor keeps its projection.  No kernel is launched."""
    scene = _scene(core, conv_gl)
    tt = getattr(core.TransformationType, trans)
    pts = np.array([[0.2, -0.1, 0.3], [-0.4, 0.25, -0.1]], np.float32)
    before = _project(scene, pts.astype(np.float64))
    target = before[0] + np.array([23.0, -11.0])
    out = core.find_transformation(pts[:1], scene, scene, core.PinUpdate(0, target.astype(np.float32)), tt)
    assert np.abs(_project(out, pts[:1].astype(np.float64))[0] - target).max() < 2e-2
    if trans == "Model":
        assert np.array_equal(out.view_matrix, scene.view_matrix)
        assert np.allclose(np.asarray(out.model_matrix)[:3, :3], np.asarray(scene.model_matrix)[:3, :3])   # pure translation
    else:
        assert np.array_equal(out.model_matrix, scene.model_matrix)
    out2 = core.find_transformation(pts, scene, scene, core.PinUpdate(0, target.astype(np.float32)), tt)
    after = _project(out2, pts.astype(np.float64))
    # two pins: rotation about the view axis through the anchor + a scale expressed as moving the anchor along its
    # view ray -- the reference's own approximation (its FIXME, pin_mode.cc:186-191): the dragged pin goes most of the
    # way, the anchor keeps its projection exactly
    assert np.linalg.norm(after[0] - target) < 0.4 * np.linalg.norm(target - before[0])
    assert np.abs(after[1] - before[1]).max() < 2e-2
    want_model, want_view = _two_pin_restatement(scene, pts.astype(np.float64), 0, target, trans)
    assert np.allclose(out2.model_matrix, want_model, atol=2e-5) and np.allclose(out2.view_matrix, want_view, atol=2e-5)
    with pytest.raises(Exception):                        # CHECK_LT(update.pin_idx, object_points.rows())
        core.find_transformation(pts, scene, scene, core.PinUpdate(2, target.astype(np.float32)), tt)
