"""GPU parity tests of the track (K10/K11) and refine (K12-K14) paths through the C ABI,
against the float32 oracle restatements (oracle/raycast.py, pnp.py, track.py, ba.py).
Tolerance: 1e-4 relative on poses / costs (BASELINE.json north_star)."""
import numpy as np
import pytest

from oracle import ba as oba
from oracle import geometry as G
from oracle import pnp as opnp
from oracle import raycast as oray
from oracle import restate, synth
from oracle import gftt as ogftt
from oracle import track as otrack
from tests import helpers as H

pytestmark = pytest.mark.gpu
F = np.float32
RTOL = 1e-4


@pytest.fixture(scope="module")
def scene(ctx_small):
    """12-frame 320x240 clip analysed by the (already parity-tested) GPU analyze path."""
    from polychase_b200 import capi
    w, h, NF = 320, 240, 12
    clip = synth.Clip(w, h, NF, seed=4)
    go = capi.default_gftt(max_corners=250)
    kps, flows = {}, {}
    ctx_small.analyze_begin(w, h, 0, NF, go)
    for k in range(NF):
        ctx_small.analyze_push(k, clip.rgb(k))
        if ctx_small.analyze_pending() >= 3:
            r = ctx_small.analyze_pop()
            kps[r["frame_id"]] = r["keypoints"]
            for (a, b, rows, idx, tgt, err) in r["pairs"]:
                flows[(a, b)] = (idx, tgt, err)
    while ctx_small.analyze_pending():
        r = ctx_small.analyze_pop()
        kps[r["frame_id"]] = r["keypoints"]
        for (a, b, rows, idx, tgt, err) in r["pairs"]:
            flows[(a, b)] = (idx, tgt, err)
    ctx_small.analyze_end()
    verts, tris = H.bumpy_mesh(clip, quads=8, amp=0.03)
    return dict(clip=clip, kps=kps, flows=flows, verts=verts, tris=tris, NF=NF, w=w, h=h)


@pytest.mark.parametrize("convention", [G.OPENCV])
def test_ray_cast_matches_bruteforce(ctx_small, scene, convention):
    clip = scene["clip"]
    cam = H.oracle_cam(clip, 3)
    model = np.eye(4, dtype=F)
    model[0, 0] = model[1, 1] = model[2, 2] = 1.0
    rng = np.random.default_rng(1)
    pos = np.stack([rng.uniform(-20, clip.width + 20, 3000), rng.uniform(-20, clip.height + 20, 3000)], 1).astype(F)
    mask = np.zeros(4, np.uint32)
    mask[0] = 0b1010_0000_0000                      # mask two triangles
    ctx_small.mesh_set(scene["verts"], scene["tris"], mask)
    hit, P, prim, uv, t = ctx_small.ray_cast(model, H.to_abi(cam), pos, True)
    o, d = oray.ray_object_space(model, cam.pose.Rt4x4(), cam.intrinsics, pos)
    eh, eP, eprim, euv, et = oray.ray_cast(scene["verts"], scene["tris"], mask, o, d, True)
    # rays grazing a shared edge may pick either neighbour / flip hit: tolerate a handful
    agree = hit == eh
    assert agree.mean() > 0.998
    both = hit & eh
    same_prim = prim[both] == eprim[both]
    assert same_prim.mean() > 0.995
    err = np.abs(P[both] - eP[both]).max(axis=1)
    assert np.percentile(err, 99.5) < 1e-4 * clip.depth
    ctx_small.mesh_set(scene["verts"], scene["tris"])


@pytest.mark.parametrize("loss,opt_f,opt_pp", [(0, False, False), (1, False, False), (2, False, False),
                                               (2, True, False), (2, True, True)])
def test_solve_pnp_matches_oracle(ctx_small, scene, loss, opt_f, opt_pp):
    from polychase_b200 import capi
    clip = scene["clip"]
    gt = H.oracle_cam(clip, 5)
    rng = np.random.default_rng(7)
    n = 1500
    X = np.stack([rng.uniform(-1.5, 1.5, n), rng.uniform(-1.0, 1.0, n), rng.uniform(-0.05, 0.05, n)], 1).astype(F)
    x = gt.intrinsics.project(gt.pose.apply(X)) + rng.normal(0, 0.3, (n, 2)).astype(F)
    x[:40] += rng.normal(0, 25, (40, 2)).astype(F)            # outliers
    init = H.perturb(gt, rng)
    opts = opnp.BundleOptions(loss_type=loss)
    ocam, ost, oinl = opnp.solve_pnp_iterative(X, x, None, init, opts, 12.0, opt_f, opt_pp)
    gcam, gst, ginl = ctx_small.solve_pnp(X, x, H.to_abi(init), capi.default_bundle(loss_type=loss), None, 12.0,
                                          opt_f, opt_pp)
    g = H.from_abi(gcam)
    dq, dt = H.pose_close(ocam, g)
    assert dq < RTOL and dt < RTOL, (dq, dt)
    assert abs(gst.cost - ost.cost) <= 1e-3 * abs(ost.cost)
    assert abs(gst.initial_cost - ost.initial_cost) <= RTOL * abs(ost.initial_cost)
    assert abs(ginl - oinl) < 2e-3
    assert abs(g.intrinsics.fy - ocam.intrinsics.fy) <= RTOL * abs(ocam.intrinsics.fy)
    assert abs(g.intrinsics.cx - ocam.intrinsics.cx) <= RTOL * abs(ocam.intrinsics.cx) + 1e-3


def test_solve_pnp_errors(ctx_small):
    from polychase_b200 import capi
    cam = capi.CameraState(100, 100, 50, 50, 1, 100, 100, 1.0)
    cam.q[:] = [1, 0, 0, 0]
    with pytest.raises(capi.PcError):          # solvers.cc:55  CHECK_GE(rows, 3)
        ctx_small.solve_pnp(np.zeros((2, 3), F), np.zeros((2, 2), F), cam)
    with pytest.raises(capi.PcError):          # solvers.cc:67-70 unknown loss
        ctx_small.solve_pnp(np.zeros((5, 3), F), np.zeros((5, 2), F), cam, capi.default_bundle(loss_type=7))


def test_track_sequence_matches_oracle(ctx_small, scene):
    """SolveFrame chained over the clip (tracker.cc:133-192): GPU poses vs the oracle's."""
    from polychase_b200 import capi
    clip, kps, flows, NF = scene["clip"], scene["kps"], scene["flows"], scene["NF"]
    model = np.eye(4, dtype=F)
    ctx_small.mesh_set(scene["verts"], scene["tris"])
    opts = opnp.BundleOptions(loss_type=opnp.CAUCHY)
    start = H.oracle_cam(clip, 0)
    want = otrack.track_sequence(kps, flows, 0, NF - 1, start, model, scene["verts"], scene["tris"], None, opts)
    traj = {0: H.to_abi(start)}
    bo = capi.default_bundle(loss_type=2)
    for f in range(1, NF):
        srcs = []
        for s in sorted(a for (a, b) in flows if b == f):
            if s in traj:
                idx, tgt, _ = flows[(s, f)]
                srcs.append((traj[s], kps[s], idx, tgt))
        cam, st, inl, m = ctx_small.track_frame(srcs, model, traj[f - 1], bo)
        traj[f] = cam
        ocam, ost, oinl, om = want[f]
        assert abs(m - om) <= max(2, 0.002 * om)
        dq, dt = H.pose_close(ocam, H.from_abi(cam))
        assert dq < RTOL and dt < RTOL, (f, dq, dt)
        assert abs(inl - oinl) < 5e-3
    # and the ground truth is recovered
    dq, dt = H.pose_close(H.oracle_cam(clip, NF - 1), H.from_abi(traj[NF - 1]))
    assert dq < 2e-3 and dt < 2e-3


@pytest.mark.parametrize("opt_f", [False, True])
def test_fused_analyze_track_chain_matches_oracle(ctx_small, scene, opt_f):
    """pc_analyze_track_begin: the forward sweep chained on the device behind the analyzer gives the
    oracle's TrackSequence poses (tracker.cc:133-213) and the same analyze rows as the plain pass."""
    from polychase_b200 import capi
    clip, kps, flows, NF = scene["clip"], scene["kps"], scene["flows"], scene["NF"]
    model = np.eye(4, dtype=F)
    opts = opnp.BundleOptions(loss_type=opnp.CAUCHY)
    start = H.oracle_cam(clip, 0)
    want = otrack.track_sequence(kps, flows, 0, NF - 1, start, model, scene["verts"], scene["tris"], None, opts,
                                 opt_f, False)
    ctx_small.mesh_set(scene["verts"], scene["tris"])
    ctx_small.analyze_begin(clip.width, clip.height, 0, NF, capi.default_gftt(max_corners=250))
    ctx_small.analyze_track_begin(model, capi.default_bundle(loss_type=2), opt_f, False)
    ctx_small.analyze_track_seed(0, H.to_abi(start))
    got = {}

    def take(r):
        got[r["frame_id"]] = r
        assert np.array_equal(r["keypoints"], kps[r["frame_id"]])
        for (a, b, rows, idx, tgt, err) in r["pairs"]:
            assert np.array_equal(idx, flows[(a, b)][0]) and np.array_equal(tgt, flows[(a, b)][1])

    for k in range(NF):
        ctx_small.analyze_push(k, clip.rgb(k))
        if ctx_small.analyze_pending() >= 3:
            take(ctx_small.analyze_pop())
    while ctx_small.analyze_pending():
        take(ctx_small.analyze_pop())
    ctx_small.analyze_end()
    assert got[0]["tracked"] == 2
    for f in range(1, NF):
        r = got[f]
        assert r["tracked"] == 1
        ocam, ost, oinl, om = want[f]
        assert abs(r["num_matches"] - om) <= max(2, 0.002 * om)
        g = H.from_abi(r["camera"])
        dq, dt = H.pose_close(ocam, g)
        # with the focal length free, depth and focal length trade off along a nearly flat valley on
        # this almost planar scene (and the chain feeds each pose to the next frame): the rotation
        # still agrees to 1e-4, translation / focal length to 1e-3
        tol = 10 * RTOL if opt_f else RTOL
        assert dq < RTOL and dt < tol, (f, dq, dt)
        assert abs(r["inlier_ratio"] - oinl) < 5e-3
        assert abs(g.intrinsics.fy - ocam.intrinsics.fy) <= tol * abs(ocam.intrinsics.fy)


def test_fused_chain_not_enough_features(ctx_small, scene):
    """A frame whose rays all miss the mesh raises the reference's error (tracker.cc:160-166) at pop."""
    from polychase_b200 import capi
    clip, NF = scene["clip"], 3
    far = scene["verts"].copy()
    far[:, 0] += 1e4                                  # mesh out of view: every ray misses
    ctx_small.mesh_set(far, scene["tris"])
    ctx_small.analyze_begin(clip.width, clip.height, 0, NF, capi.default_gftt(max_corners=250))
    ctx_small.analyze_track_begin(np.eye(4, dtype=F), capi.default_bundle())
    ctx_small.analyze_track_seed(0, H.to_abi(H.oracle_cam(clip, 0)))
    ctx_small.analyze_push(0, clip.rgb(0))
    ctx_small.analyze_push(1, clip.rgb(1))
    assert ctx_small.analyze_pop()["tracked"] == 2
    with pytest.raises(capi.PcError) as e:
        ctx_small.analyze_pop()
    assert e.value.code == -7 and "Could not track to frame: 1" in str(e.value)
    ctx_small.analyze_end()
    ctx_small.mesh_set(scene["verts"], scene["tris"])


def test_track_not_enough_features(ctx_small, scene):
    from polychase_b200 import capi
    ctx_small.mesh_set(scene["verts"], scene["tris"])
    cam = H.to_abi(H.oracle_cam(scene["clip"], 0))
    with pytest.raises(capi.PcError) as e:
        ctx_small.track_frame([(cam, np.zeros((2, 2), F), np.array([0, 1], np.uint32), np.zeros((2, 2), F))],
                              np.eye(4, dtype=F), cam)
    assert e.value.code == -7


def _ba_setup(scene, rng, opt_f=False, opt_pp=False, perturb=True):
    clip, kps, flows, NF = scene["clip"], scene["kps"], scene["flows"], scene["NF"]
    model = np.eye(4, dtype=F)
    traj = [H.oracle_cam(clip, k) for k in range(NF)]
    if perturb:
        for k in range(1, NF - 1):
            traj[k] = H.perturb(traj[k], rng, rot_deg=0.05, trans=0.004)
    edges = [oba.Edge(a, b, flows[(a, b)][0], flows[(a, b)][1]) for (a, b) in sorted(flows) if len(flows[(a, b)][0])]
    prob = oba.RefineProblem([kps[k] for k in range(NF)], edges, scene["verts"], scene["tris"], None, model,
                             opt_f, opt_pp, traj[0].intrinsics.bounds())
    abi_edges = [(e.src, e.tgt, e.src_kps_indices, e.tgt_kps) for e in edges]
    return model, traj, prob, abi_edges


@pytest.mark.parametrize("opt_f,opt_pp,loss", [(False, False, 2), (True, True, 2), (False, False, 1)])
def test_ba_cost_and_normal_equations(ctx_small, scene, opt_f, opt_pp, loss):
    from polychase_b200 import capi
    rng = np.random.default_rng(3)
    model, traj, prob, abi_edges = _ba_setup(scene, rng, opt_f, opt_pp)
    ctx_small.mesh_set(scene["verts"], scene["tris"])
    ctx_small.ba_load([scene["kps"][k] for k in range(scene["NF"])], abi_edges, model, opt_f, opt_pp)
    bo = capi.default_bundle(loss_type=loss)
    lo = opnp.Loss(loss, 1.0)
    atraj = [H.to_abi(c) for c in traj]
    want_cost = prob.total_cost(traj, lo)
    got_cost = ctx_small.ba_cost(atraj, bo)
    assert abs(got_cost - want_cost) <= 2e-4 * abs(want_cost)
    A, g = prob.normal_equations(traj, lo)
    band, jtr = ctx_small.ba_normal_equations(atraj, bo)
    Ag = capi.band_to_dense(band)
    scale = np.abs(A).max()
    assert np.abs(Ag - np.tril(A)).max() <= 5e-4 * scale
    assert np.abs(jtr - g).max() <= 5e-4 * np.abs(g).max()
    # first / last frame are ground truth: their blocks are empty (refiner.cc:611-612)
    p = prob.p
    assert np.all(Ag[:p, :] == 0) and np.all(Ag[-p:, :] == 0)


@pytest.mark.parametrize("opt_f,opt_pp", [(False, False), (True, False)])
def test_ba_solve_matches_oracle(ctx_small, scene, opt_f, opt_pp):
    from polychase_b200 import capi
    rng = np.random.default_rng(5)
    model, traj, prob, abi_edges = _ba_setup(scene, rng, opt_f, opt_pp)
    ctx_small.mesh_set(scene["verts"], scene["tris"])
    ctx_small.ba_load([scene["kps"][k] for k in range(scene["NF"])], abi_edges, model, opt_f, opt_pp)
    opts = opnp.BundleOptions(loss_type=opnp.CAUCHY, max_iterations=30)
    want, wst = oba.refine_trajectory(prob, traj, opts)
    seen = []
    got, gst = ctx_small.ba_solve([H.to_abi(c) for c in traj], capi.default_bundle(loss_type=2, max_iterations=30),
                                  callback=lambda s: seen.append(s.cost) or True)
    assert len(seen) >= 1
    assert abs(gst.initial_cost - wst.initial_cost) <= 2e-4 * abs(wst.initial_cost)
    assert gst.cost <= gst.initial_cost
    assert abs(gst.cost - wst.cost) <= 2e-3 * abs(wst.cost)
    for k in range(len(traj)):
        dq, dt = H.pose_close(want[k], H.from_abi(got[k]))
        assert dq < 2e-4 and dt < 2e-4, (k, dq, dt)
    # end frames untouched
    for k in (0, len(traj) - 1):
        assert list(got[k].q) == [float(v) for v in traj[k].pose.q]
