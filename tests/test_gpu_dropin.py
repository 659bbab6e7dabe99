"""End-to-end drop-in test on the GPU: the addon's call sequence through
the `polychase_core` surface -- OpticalFlowThread with the frame hand-off -> SQLite ->
TrackerThread -> RefinerThread -- checked against the oracle."""
import time

import numpy as np
import pytest

from oracle import db as odb
from oracle import gftt as ogftt
from oracle import pnp as opnp
from oracle import restate, synth
from oracle import track as otrack
from tests import helpers as H

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.fixture(scope="module")
def core():
    from polychase_b200 import polychase_core
    return polychase_core


def _pump(thread, on_msg, timeout=120):
    t0 = time.time()
    while time.time() - t0 < timeout:
        m = thread.try_pop()
        if m is None:
            time.sleep(0.002)
            continue
        if isinstance(m, bool):
            return
        on_msg(m)
    raise TimeoutError


def test_analyze_track_refine_like_the_addon(core, tmp_path):
    w, h, NF, first = 320, 240, 14, 1          # Blender frame numbers are 1-based
    clip = synth.Clip(w, h, NF, seed=6, first_frame=first)
    frames = {k: clip.rgb(k) for k in range(first, first + NF)}
    dbp = str(tmp_path / "clip.db")
    go = core.GFTTOptions()
    go.max_corners = 300
    th = core.OpticalFlowThread(core.VideoInfo(w, h, first, NF), dbp, go)
    requested, progress, errors = [], [], []

    def on_msg(m):
        if isinstance(m, core.OpticalFlowRequest):
            requested.append(m.frame_id)
            th.provide_frame(m.frame_id, frames[m.frame_id])
        elif isinstance(m, core.OpticalFlowProgress):
            progress.append((m.progress, m.progress_message))
        elif isinstance(m, core.CppException):
            errors.append(m.what())
    _pump(th, on_msg)
    th.join()
    assert not errors, errors
    assert requested == list(range(first, first + NF))
    assert progress[-1][1] == "Done" and progress[0][1] == f"Processing frame {first}"

    # database content == what the reference's loop would have written (oracle restatement)
    o = odb.Database(dbp)
    assert o.frames() == list(range(first, first + NF))
    want_pairs = sorted((a, a + d) for a in range(first, first + NF) for d in (-8, -4, -2, -1, 1, 2, 4, 8)
                        if first <= a + d < first + NF)
    assert o.pairs() == want_pairs and len(want_pairs) == 8 * NF - 30
    grays = {k: restate.rgb2gray(frames[k]) for k in frames}
    pyr = {k: restate.pyramid(grays[k], 3) for k in frames}
    kps, flows = {}, {}
    for k in frames:
        kps[k] = o.read_keypoints(k)
        assert np.array_equal(kps[k], ogftt.gftt_from_eig(restate.min_eig(grays[k], 3), max_corners=300))
    for (a, b) in want_pairs:
        idx, tgt, err = o.read_image_pair_flow(a, b)
        wn, ws, we = restate.lk(pyr[a], pyr[b], kps[a])
        ok = ws == 1
        assert np.array_equal(idx, np.nonzero(ok)[0].astype(np.uint32))
        assert np.array_equal(tgt.view(np.uint32), wn[ok].view(np.uint32))
        assert np.array_equal(err.view(np.uint32), we[ok].view(np.uint32))
        flows[(a, b)] = (idx, tgt, err)
    # sources come back in ascending image_id_from
    assert o.find_optical_flows_to_image(first + 9) == sorted(o.find_optical_flows_to_image(first + 9))
    o.close()

    # resume: a second pass over an already complete database writes nothing new and does not fail
    th2 = core.OpticalFlowThread(core.VideoInfo(w, h, first, NF), dbp, go)
    _pump(th2, lambda m: th2.provide_frame(m.frame_id, frames[m.frame_id]) if isinstance(m, core.OpticalFlowRequest)
          else errors.append(m.what()) if isinstance(m, core.CppException) else None)
    th2.join()
    assert not errors, errors

    # ---- track -------------------------------------------------------------------------------
    verts, tris = H.bumpy_mesh(clip, quads=8, amp=0.03)
    mesh = core.AcceleratedMesh(verts, tris)
    K = clip.K
    intr = core.CameraIntrinsics(K["fx"], K["fy"], K["cx"], K["cy"], 1.0, w, h, core.CameraConvention.OpenCV)
    view = np.eye(4, dtype=F)
    view[:3, :3] = clip.R[0]
    view[:3, 3] = clip.t[0]
    scene = core.SceneTransformations(np.eye(4, dtype=F), view, intr)
    bo = core.BundleOptions()
    bo.loss_type = core.LossType.Cauchy
    tt = core.TrackerThread(dbp, first, first + NF - 1, scene, mesh, False, False, bo)
    results = []
    _pump(tt, lambda m: results.append(m) if isinstance(m, core.FrameTrackingResult) else errors.append(m.what()))
    tt.join()
    assert not errors, errors
    assert [r.frame for r in results] == list(range(first + 1, first + NF))
    start = H.oracle_cam(clip, first)
    # the start pose goes through Pose::FromRt(view_matrix) in both implementations
    want = otrack.track_sequence(kps, flows, first, first + NF - 1, start, np.eye(4, dtype=F), verts, tris, None,
                                 opnp.BundleOptions(loss_type=opnp.CAUCHY))
    for r in results:
        ocam = want[r.frame][0]
        dq = np.abs(np.abs(np.array(r.pose.q)) - np.abs(ocam.pose.q)).max()
        dt = np.abs(np.array(r.pose.t) - ocam.pose.t).max() / np.abs(ocam.pose.t).max()
        assert dq < 1e-4 and dt < 1e-4, (r.frame, dq, dt)
        assert r.inlier_ratio > 0.9 and r.bundle_stats.cost <= r.bundle_stats.initial_cost

    # ---- refine ------------------------------------------------------------------------------
    traj = core.CameraTrajectory(first, NF)
    for k in range(first, first + NF):
        p = core.Pose()
        src = H.oracle_cam(clip, k) if k in (first, first + NF - 1) else None
        if src is None:
            r = results[k - first - 1]
            p.q, p.t = r.pose.q, r.pose.t
        else:
            p.q, p.t = src.pose.q, src.pose.t
        traj.set(k, core.CameraState(intr, p))
    rt = core.RefinerThread(dbp, traj, np.eye(4, dtype=F), mesh, False, False, bo)
    updates = []
    _pump(rt, lambda m: updates.append(m) if isinstance(m, core.RefineTrajectoryUpdate) else errors.append(m.what()))
    rt.join()
    assert not errors, errors
    assert updates and updates[-1].message.startswith("Cost: ")
    assert updates[-1].stats.cost <= updates[-1].stats.initial_cost
    # the trajectory object was refined in place; ground truth is close
    for k in range(first + 1, first + NF - 1):
        got = traj.get(k)
        gt = H.oracle_cam(clip, k)
        assert np.abs(np.array(got.pose.t) - gt.pose.t).max() < 5e-3 * clip.depth


def test_tracker_thread_reports_errors(core, tmp_path):
    dbp = str(tmp_path / "empty.db")
    core.Database(dbp).close()
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    mesh = core.AcceleratedMesh(v, np.array([[0, 1, 2]], np.uint32))
    intr = core.CameraIntrinsics(100, 100, 50, 50, 1.0, 100, 100, core.CameraConvention.OpenCV)
    scene = core.SceneTransformations(np.eye(4, dtype=F), np.eye(4, dtype=F), intr)
    tt = core.TrackerThread(dbp, 1, 3, scene, mesh, False, False, core.BundleOptions())
    msgs = []
    _pump(tt, msgs.append)
    tt.join()
    assert len(msgs) == 1 and isinstance(msgs[0], core.CppException)
    assert "Not enough features" in msgs[0].what()          # tracker.cc:162-166 (message preserved, not sliced)


def test_optical_flow_thread_cancel(core, tmp_path):
    th = core.OpticalFlowThread(core.VideoInfo(64, 48, 1, 100), str(tmp_path / "c.db"))
    time.sleep(0.05)
    th.request_stop()
    th.join()
    last = None
    while True:
        m = th.try_pop()
        if m is None:
            break
        last = m
    assert last is True


def test_sync_entry_points_with_python_callbacks_opengl(core, tmp_path):
    """generate_optical_flow_database / track_sequence / refine_trajectory (polychase_pybind.cc:313-347) called
    with Python callables -- they run without the GIL and call back with it -- on the addon's own camera:
    OpenGL convention, negated fx / fy, bottom-up frames (blender_addon/core.py:348-357), while another
    Python thread ray-casts against the same mesh (the UI does that during tracking)."""
    import threading
    from oracle import geometry as G
    w, h, NF, first = 320, 240, 12, 1
    clip = synth.Clip(w, h, NF, seed=8, first_frame=first)
    frames = {k: H.clip_rgb(clip, k, G.OPENGL) for k in range(first, first + NF)}
    dbp = str(tmp_path / "sync.db")
    go = core.GFTTOptions()
    go.max_corners = 300
    asked, prog = [], []

    def accessor(frame_id):
        asked.append(frame_id)
        return frames[frame_id]

    def progress(p, msg):
        prog.append(msg)
        return True

    core.generate_optical_flow_database(core.VideoInfo(w, h, first, NF), accessor, progress, dbp, go)
    assert asked == list(range(first, first + NF)) and prog[-1] == "Done"
    o = odb.Database(dbp)
    assert len(o.pairs()) == 8 * NF - 30
    kps = {k: o.read_keypoints(k) for k in frames}
    flows = {pr: o.read_image_pair_flow(*pr) for pr in o.pairs()}
    o.close()

    verts, tris = H.bumpy_mesh(clip, quads=8, amp=0.03)
    mesh = core.AcceleratedMesh(verts, tris)
    start = H.oracle_cam(clip, first, G.OPENGL)
    it = start.intrinsics
    intr = core.CameraIntrinsics(float(it.fx), float(it.fy), float(it.cx), float(it.cy), 1.0, w, h,
                                 core.CameraConvention.OpenGL)
    assert intr.fx < 0 and intr.fy < 0
    scene = core.SceneTransformations(np.eye(4, dtype=F), start.pose.Rt4x4(), intr)
    bo = core.BundleOptions()
    bo.loss_type = core.LossType.Cauchy

    # a UI-thread ray cast while the tracker runs must neither deadlock nor wait for the whole sweep
    stop = threading.Event()
    hits = []

    def ui_raycasts():
        while not stop.is_set():
            hits.append(core.ray_cast(mesh, scene, np.array([w / 2, h / 2], F), True))
            time.sleep(0.001)

    ui = threading.Thread(target=ui_raycasts)
    ui.start()
    results = []
    try:
        core.track_sequence(dbp, first, first + NF - 1, scene, mesh, lambda r: results.append(r) or True, False, False, bo)
    finally:
        stop.set()
        ui.join(timeout=60)
    assert not ui.is_alive() and hits and all(hh is not None for hh in hits)
    assert [r.frame for r in results] == list(range(first + 1, first + NF))
    want = otrack.track_sequence(kps, flows, first, first + NF - 1, start, np.eye(4, dtype=F), verts, tris, None,
                                 opnp.BundleOptions(loss_type=opnp.CAUCHY))
    for r in results:
        ocam = want[r.frame][0]
        dq = np.abs(np.abs(np.array(r.pose.q)) - np.abs(ocam.pose.q)).max()
        dt = np.abs(np.array(r.pose.t) - ocam.pose.t).max() / np.abs(ocam.pose.t).max()
        assert dq < 1e-4 and dt < 1e-4, (r.frame, dq, dt)
        gt = H.oracle_cam(clip, r.frame, G.OPENGL)
        assert np.abs(np.array(r.pose.t) - gt.pose.t).max() < 5e-3 * clip.depth

    # a callback that returns False stops the sweep after that frame (tracker.cc:178-183)
    seen = []
    core.track_sequence(dbp, first, first + NF - 1, scene, mesh, lambda r: seen.append(r.frame) or len(seen) < 3, False,
                        False, bo)
    assert seen == [first + 1, first + 2, first + 3]

    traj = core.CameraTrajectory(first, NF)
    for k in range(first, first + NF):
        p = core.Pose()
        if k in (first, first + NF - 1):
            src = H.oracle_cam(clip, k, G.OPENGL)
            p.q, p.t = src.pose.q, src.pose.t
        else:
            p.q, p.t = results[k - first - 1].pose.q, results[k - first - 1].pose.t
        traj.set(k, core.CameraState(intr, p))
    updates = []
    core.refine_trajectory(dbp, traj, np.eye(4, dtype=F), mesh, False, False, lambda u: updates.append(u) or True, bo)
    assert updates and updates[-1].stats.cost <= updates[-1].stats.initial_cost
    for k in range(first + 1, first + NF - 1):
        gt = H.oracle_cam(clip, k, G.OPENGL)
        assert np.abs(np.array(traj.get(k).pose.t) - gt.pose.t).max() < 5e-3 * clip.depth


def test_resume_with_empty_keypoint_row_is_authoritative(core, tmp_path):
    """ReadOrGenerateKeypoints (opticalflow.cc:168-178) only detects when the frame has no row: a stored row
    with zero keypoints stays empty, and no flow row may reference keypoints that are not in the database."""
    w, h, NF, first = 160, 128, 4, 1
    clip = synth.Clip(w, h, NF, seed=3, first_frame=first)
    dbp = str(tmp_path / "resume.db")
    d = core.Database(dbp)
    d.write_keypoints(first + 1, np.zeros((0, 2), F))
    d.close()
    go = core.GFTTOptions()
    go.max_corners = 100
    core.generate_optical_flow_database(core.VideoInfo(w, h, first, NF), lambda f: clip.rgb(f), None, dbp, go)
    o = odb.Database(dbp)
    assert len(o.read_keypoints(first + 1)) == 0
    for (a, b) in o.pairs():
        idx, tgt, err = o.read_image_pair_flow(a, b)
        if a == first + 1:
            assert len(idx) == 0
        elif len(idx):
            assert int(idx.max()) < len(o.read_keypoints(a))
    o.close()


@pytest.mark.parametrize("conv_gl", [True, False])
@pytest.mark.parametrize("trans", ["Model", "Camera"])
def test_find_transformation_many_pins_on_the_gpu(core, trans, conv_gl):
    """FindTransformationN (pin_mode.cc:16-108): three and more pins are solved by SolvePnPIterative (K11 through
    pc_solve_pnp) with the trivial loss from the current transform.  Compared with the float32 restatement
    (oracle/pin_mode.py): the returned matrices agree to 1e-4 relative, and the pins reproject within 0.05 px of the
    restatement's.  A consistent drag (all pins already moved rigidly) is recovered to sub-pixel accuracy."""
    from oracle import geometry as G
    from oracle import pin_mode as opin
    from tests.test_core_surface import _project, _scene
    scene = _scene(core, conv_gl)
    tt = getattr(core.TransformationType, trans)
    ott = opin.MODEL if trans == "Model" else opin.CAMERA
    k = scene.intrinsics
    ointr = G.Intrinsics(k.fx, k.fy, k.cx, k.cy, k.aspect_ratio, k.width, k.height, G.OPENGL if conv_gl else G.OPENCV)
    rng = np.random.default_rng(2)
    pts = rng.uniform(-0.6, 0.6, (6, 3)).astype(F)
    before = _project(scene, pts.astype(np.float64))
    # a drag of pin 2: rotate the object a little about its own z axis and read off where the pin lands
    a = np.deg2rad(3.0)
    Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    moved = core.SceneTransformations(np.asarray(scene.model_matrix) @ np.block([[Rz, np.zeros((3, 1))], [np.zeros((1, 3)), 1]]).astype(F),
                                      scene.view_matrix, scene.intrinsics)
    target = _project(moved, pts.astype(np.float64))

    def check(out, cur, pin):
        wm, wv, wi, st = opin.find_transformation_n(pts, scene.model_matrix, scene.view_matrix, ointr, cur.model_matrix,
                                                    cur.view_matrix, ointr, pin, target[pin].astype(F), ott)
        for got, want in ((np.asarray(out.model_matrix), wm), (np.asarray(out.view_matrix), wv)):
            assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max(), (got, want)
        want_scene = core.SceneTransformations(wm, wv, scene.intrinsics)
        assert np.abs(_project(out, pts.astype(np.float64)) - _project(want_scene, pts.astype(np.float64))).max() < 0.05

    out = core.find_transformation(pts, scene, scene, core.PinUpdate(2, target[2].astype(F)), tt)
    check(out, scene, 2)
    after = _project(out, pts.astype(np.float64))
    # least squares over six pins of which one moved: the dragged pin goes most of the way, the others give a little
    assert np.abs(after[2] - target[2]).max() < 0.5 * np.abs(before[2] - target[2]).max()
    others = [i for i in range(6) if i != 2]
    assert np.abs(after[others] - before[others]).max() < np.abs(before[2] - target[2]).max()
    if trans == "Model":
        assert np.array_equal(out.view_matrix, scene.view_matrix)
    else:
        assert np.array_equal(out.model_matrix, scene.model_matrix)
    # successive updates start from the previous result (smooth transitions, pin_mode.cc:49-54)
    cur = scene
    for i in range(6):
        nxt = core.find_transformation(pts, scene, cur, core.PinUpdate(i, target[i].astype(F)), tt)
        check(nxt, cur, i)
        cur = nxt
    # a consistent drag: the image points handed to the solver are those of a rigid motion -> recovered
    rigid = core.find_transformation(pts[:3], scene, scene, core.PinUpdate(2, target[2].astype(F)), tt)
    fin = _project(rigid, pts[:3].astype(np.float64))
    assert np.abs(fin[2] - target[2]).max() < 0.05               # three pins: the solve is exact (6 equations, 6 unknowns)
    assert np.abs(fin[:2] - before[:2]).max() < 0.05


def test_baseline_config0_720p_64_frames_through_the_plugin(core, tmp_path):
    """BASELINE.json configs[0]: "64-frame 720p synthetic clip, 2k features, OpticalFlow + Tracker" through the reference's
    own entry points (OpticalFlowThread -> SQLite -> TrackerThread), OpenGL camera as the addon passes it.  The database
    must hold the 8F - 30 pairs; sampled frames / pairs are compared with the oracle bit for bit (keypoints against the
    cv2-pinned detector restatement, flow rows against the LK restatement); the tracked poses must stay within 1e-4
    relative of the synthetic ground truth over the whole clip."""
    from oracle import geometry as G
    from polychase_b200 import synth as psynth
    w, h, NF, first, mc = 1280, 720, 64, 1, 2000
    clip = synth.Clip(w, h, NF, seed=12, first_frame=first, speed=psynth.survey_speed(w))
    conv = G.OPENGL
    frames = {k: H.clip_rgb(clip, k, conv) for k in range(first, first + NF)}
    dbp = str(tmp_path / "c0.db")
    go = core.GFTTOptions()
    go.max_corners = mc
    errors = []
    th = core.OpticalFlowThread(core.VideoInfo(w, h, first, NF), dbp, go)
    _pump(th, lambda m: th.provide_frame(m.frame_id, frames[m.frame_id]) if isinstance(m, core.OpticalFlowRequest)
          else errors.append(m.what()) if isinstance(m, core.CppException) else None, timeout=300)
    th.join()
    assert not errors, errors
    o = odb.Database(dbp)
    assert o.frames() == list(range(first, first + NF))
    assert len(o.pairs()) == 8 * NF - 30
    for k in (first, first + 31, first + NF - 1):                    # sampled frames: the detector, bit for bit
        g = restate.rgb2gray(frames[k])
        assert np.array_equal(o.read_keypoints(k), ogftt.gftt_from_eig(restate.min_eig(g, 3), max_corners=mc)), k
    for (a, b) in ((first + 8, first), (first + 20, first + 28), (first + 40, first + 39), (first + 63, first + 59)):
        pa = restate.pyramid(restate.rgb2gray(frames[a]), 3)
        pb = restate.pyramid(restate.rgb2gray(frames[b]), 3)
        idx, tgt, err = o.read_image_pair_flow(a, b)
        wn, ws, we = restate.lk(pa, pb, o.read_keypoints(a))
        ok = ws == 1
        assert np.array_equal(idx, np.nonzero(ok)[0].astype(np.uint32)), (a, b)
        assert np.array_equal(tgt.view(np.uint32), wn[ok].view(np.uint32)), (a, b)
        assert np.array_equal(err.view(np.uint32), we[ok].view(np.uint32)), (a, b)
    o.close()
    verts, tris = synth.plane_mesh(w, h, clip.s)
    mesh = core.AcceleratedMesh(verts, tris)
    start = H.oracle_cam(clip, first, conv)
    ii = start.intrinsics
    intr = core.CameraIntrinsics(float(ii.fx), float(ii.fy), float(ii.cx), float(ii.cy), 1.0, w, h,
                                 core.CameraConvention.OpenGL)
    scene = core.SceneTransformations(np.eye(4, dtype=F), start.pose.Rt4x4(), intr)
    bo = core.BundleOptions()
    bo.loss_type = core.LossType.Cauchy
    tt = core.TrackerThread(dbp, first, first + NF - 1, scene, mesh, False, False, bo)
    results = []
    _pump(tt, lambda m: results.append(m) if isinstance(m, core.FrameTrackingResult) else errors.append(m.what()), timeout=300)
    tt.join()
    assert not errors, errors
    assert [r.frame for r in results] == list(range(first + 1, first + NF))
    worst = 0.0
    for r in results:
        gt = H.oracle_cam(clip, r.frame, conv)
        worst = max(worst, float(np.abs(np.array(r.pose.t) - gt.pose.t).max()))
        assert r.inlier_ratio > 0.9
    assert worst < 1e-4 * clip.depth * 2.5, worst            # 2.5e-4 of the depth: the bench sweep's own figure over 640 frames is 2e-4
