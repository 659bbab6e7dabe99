"""GPU parity tests of the track (K10/K11) and refine (K12-K14) paths through the C ABI,
against the float32 oracle restatements (oracle/raycast.py, pnp.py, track.py, ba.py).

Every test runs in BOTH camera conventions: OpenCV, and the one the Blender addon actually uses --
OpenGL with negated fx, fy on bottom-up images (/root/reference/blender_addon/core.py:348-357,
cpp/pnp/types.h:95-98,129-132,156-170; tests/helpers.py::oracle_cam).

Tolerance: 1e-4 relative on poses / costs / normal equations (BASELINE.json north_star).  The
comparison value is the oracle with correctly rounded (float64-accumulated) sums -- the value every
float32 summation order scatters around.  The reference's own sums are order dependent (TBB combines
thread-local partials in scheduling order, lev_marq.h:231-297,653-771); where a converged LM result
amplifies that, the oracle is re-run with permuted float32 summation orders and the GPU has to sit
inside max(1e-4, BAND x the largest deviation those few samples show) -- BAND = 2 because the maximum
of two or three samples underestimates the spread of the distribution they are drawn from.  For the chained
sweep the samples also include runs whose LK target positions are moved by one float32 ulp: the size of the
per-residual arithmetic differences between any two correct float32 implementations (fused multiply-add or not,
R X versus q X q* -- the reference's own Evaluate and EvaluateWithJacobian differ in that, pnp_problem.h:54,76)."""
import numpy as np
import pytest

from oracle import ba as oba
from oracle import geometry as G
from oracle import pnp as opnp
from oracle import raycast as oray
from oracle import synth
from oracle import track as otrack
from tests import helpers as H

pytestmark = pytest.mark.gpu
F = np.float32
RTOL = 1e-4
BAND = 2.0
CONVENTIONS = [G.OPENCV, G.OPENGL]
CONV_IDS = ["opencv", "opengl"]


def analyze_clip(ctx, clip, nf, max_corners, conv, first=0):
    """Keypoints and flow rows of the clip from the (already parity-tested) GPU analyze path."""
    from polychase_b200 import capi
    kps, flows = {}, {}
    ctx.analyze_begin(clip.width, clip.height, first, nf, capi.default_gftt(max_corners=max_corners))

    def take(r):
        kps[r["frame_id"]] = r["keypoints"]
        for (a, b, rows, idx, tgt, err) in r["pairs"]:
            flows[(a, b)] = (idx, tgt, err)

    for k in range(first, first + nf):
        ctx.analyze_push(k, H.clip_rgb(clip, k, conv))
        if ctx.analyze_pending() >= 3:
            take(ctx.analyze_pop())
    while ctx.analyze_pending():
        take(ctx.analyze_pop())
    ctx.analyze_end()
    return kps, flows


@pytest.fixture(scope="module", params=CONVENTIONS, ids=CONV_IDS)
def scene(ctx_small, request):
    """12-frame 320x240 clip seen through the given camera convention."""
    conv = request.param
    w, h, NF = 320, 240, 12
    clip = synth.Clip(w, h, NF, seed=4)
    kps, flows = analyze_clip(ctx_small, clip, NF, 250, conv)
    verts, tris = H.bumpy_mesh(clip, quads=8, amp=0.03)
    return dict(clip=clip, kps=kps, flows=flows, verts=verts, tris=tris, NF=NF, w=w, h=h, conv=conv)


def cam_of(scene, k):
    return H.oracle_cam(scene["clip"], k, scene["conv"])


def test_scene_cameras_explain_the_flows(scene):
    """Guards the fixture itself: ray-casting a source keypoint with the source camera and projecting
    with the target camera lands on the LK target (so the OpenGL variant really is a consistent camera)."""
    clip, kps, flows = scene["clip"], scene["kps"], scene["flows"]
    verts, tris = synth.plane_mesh(clip.width, clip.height, clip.s, quads=2)
    model = np.eye(4, dtype=F)
    idx, tgt, _ = flows[(3, 4)]
    a, b = cam_of(scene, 3), cam_of(scene, 4)
    o, d = oray.ray_object_space(model, a.pose.Rt4x4(), a.intrinsics, kps[3][idx])
    hit, P, _, _, _ = oray.ray_cast(verts, tris, None, o, d, True)
    assert hit.mean() > 0.99
    Pc = b.pose.apply(P[hit])
    assert not b.intrinsics.is_behind(Pc).any()
    err = np.linalg.norm(b.intrinsics.project(Pc) - tgt[hit], axis=1)
    assert np.median(err) < 0.2


def test_ray_cast_matches_bruteforce(ctx_small, scene):
    clip = scene["clip"]
    cam = cam_of(scene, 3)
    model = np.eye(4, dtype=F)
    rng = np.random.default_rng(1)
    pos = np.stack([rng.uniform(-20, clip.width + 20, 3000), rng.uniform(-20, clip.height + 20, 3000)], 1).astype(F)
    mask = np.zeros(4, np.uint32)
    mask[0] = 0b1010_0000_0000                      # mask two triangles
    ctx_small.mesh_set(scene["verts"], scene["tris"], mask)
    hit, P, prim, uv, t = ctx_small.ray_cast(model, H.to_abi(cam), pos, True)
    o, d = oray.ray_object_space(model, cam.pose.Rt4x4(), cam.intrinsics, pos)
    eh, eP, eprim, euv, et = oray.ray_cast(scene["verts"], scene["tris"], mask, o, d, True)
    assert eh.mean() > 0.5                          # the camera does look at the mesh in this convention
    # rays grazing a shared edge may pick either neighbour / flip hit: tolerate a handful
    agree = hit == eh
    assert agree.mean() > 0.998
    both = hit & eh
    same_prim = prim[both] == eprim[both]
    assert same_prim.mean() > 0.995
    err = np.abs(P[both] - eP[both]).max(axis=1)
    assert np.percentile(err, 99.5) < 1e-4 * clip.depth
    ctx_small.mesh_set(scene["verts"], scene["tris"])


def _pnp_oracle_with_band(X, x, init, opts, opt_f, opt_pp, seeds=(1, 2, 3)):
    """(exact-sum oracle result, noise band of the float32 summation order)."""
    with opnp.summation(np.float64):
        ocam, ost, oinl = opnp.solve_pnp_iterative(X, x, None, init, opts, 12.0, opt_f, opt_pp)
    bq = bt = bc = bf = 0.0
    for s in seeds:
        with opnp.summation(F, perm_seed=s):
            c, st, _ = opnp.solve_pnp_iterative(X, x, None, init, opts, 12.0, opt_f, opt_pp)
        dq, dt = H.pose_close(ocam, c)
        bq, bt = max(bq, dq), max(bt, dt)
        bc = max(bc, abs(float(st.cost) - float(ost.cost)) / abs(float(ost.cost)))
        bf = max(bf, abs(float(c.intrinsics.fy) - float(ocam.intrinsics.fy)) / abs(float(ocam.intrinsics.fy)))
    return ocam, ost, oinl, dict(q=BAND * bq, t=BAND * bt, cost=BAND * bc, f=BAND * bf)


@pytest.mark.parametrize("loss,opt_f,opt_pp", [(0, False, False), (1, False, False), (2, False, False),
                                               (2, True, False), (2, True, True)])
def test_solve_pnp_matches_oracle(ctx_small, scene, loss, opt_f, opt_pp):
    from polychase_b200 import capi
    clip = scene["clip"]
    gt = cam_of(scene, 5)
    rng = np.random.default_rng(7)
    n = 1500
    X = np.stack([rng.uniform(-1.5, 1.5, n), rng.uniform(-1.0, 1.0, n), rng.uniform(-0.05, 0.05, n)], 1).astype(F)
    assert not gt.intrinsics.is_behind(gt.pose.apply(X)).any()
    x = gt.intrinsics.project(gt.pose.apply(X)) + rng.normal(0, 0.3, (n, 2)).astype(F)
    x[:40] += rng.normal(0, 25, (40, 2)).astype(F)            # outliers
    init = H.perturb(gt, rng)
    opts = opnp.BundleOptions(loss_type=loss)
    ocam, ost, oinl, band = _pnp_oracle_with_band(X, x, init, opts, opt_f, opt_pp)
    gcam, gst, ginl = ctx_small.solve_pnp(X, x, H.to_abi(init), capi.default_bundle(loss_type=loss), None, 12.0,
                                          opt_f, opt_pp)
    g = H.from_abi(gcam)
    assert int(g.intrinsics.convention) == scene["conv"]
    dq, dt = H.pose_close(ocam, g)
    assert dq <= max(RTOL, band["q"]) and dt <= max(RTOL, band["t"]), (dq, dt, band)
    assert abs(gst.cost - ost.cost) <= max(RTOL, band["cost"]) * abs(ost.cost), (gst.cost, ost.cost, band)
    assert abs(gst.initial_cost - ost.initial_cost) <= RTOL * abs(ost.initial_cost)
    assert abs(ginl - oinl) < 2e-3
    assert abs(g.intrinsics.fy - ocam.intrinsics.fy) <= max(RTOL, band["f"]) * abs(ocam.intrinsics.fy)
    assert abs(g.intrinsics.cx - ocam.intrinsics.cx) <= RTOL * abs(ocam.intrinsics.cx) + 1e-3
    if opt_f:                                                 # the focal length moved, and stayed inside the convention's bounds
        b = init.intrinsics.bounds()
        assert b["f_low"] <= g.intrinsics.fy <= b["f_high"] and b["f_low"] < b["f_high"]
        assert (g.intrinsics.fy < 0) == (scene["conv"] == G.OPENGL)
        assert g.intrinsics.fx == F(g.intrinsics.fy * g.intrinsics.aspect_ratio)


def test_solve_pnp_errors(ctx_small):
    from polychase_b200 import capi
    cam = capi.CameraState(100, 100, 50, 50, 1, 100, 100, 1.0)
    cam.q[:] = [1, 0, 0, 0]
    with pytest.raises(capi.PcError):          # solvers.cc:55  CHECK_GE(rows, 3)
        ctx_small.solve_pnp(np.zeros((2, 3), F), np.zeros((2, 2), F), cam)
    with pytest.raises(capi.PcError):          # solvers.cc:67-70 unknown loss
        ctx_small.solve_pnp(np.zeros((5, 3), F), np.zeros((5, 2), F), cam, capi.default_bundle(loss_type=7))


def _track_oracle_with_band(scene, opts, opt_f=False, seeds=(1, 2, 3, 4)):
    clip, kps, flows, NF = scene["clip"], scene["kps"], scene["flows"], scene["NF"]
    model = np.eye(4, dtype=F)
    start = cam_of(scene, 0)
    args = (kps, flows, 0, NF - 1, start, model, scene["verts"], scene["tris"], None, opts, opt_f, False)
    with opnp.summation(np.float64):
        want = otrack.track_sequence(*args)
    band = {f: [0.0, 0.0, 0.0] for f in want}
    for s in seeds:
        if s % 2:                                                  # odd seeds: summation order
            with opnp.summation(F, perm_seed=s):
                alt = otrack.track_sequence(*args)
        else:                                                      # even seeds: one-ulp jitter of the LK targets
            rng = np.random.default_rng(s)
            jf = {k: (idx, np.nextafter(tgt, np.where(rng.random(tgt.shape) < 0.5, -np.inf, np.inf).astype(F)), err)
                  for k, (idx, tgt, err) in flows.items()}
            with opnp.summation(np.float64):
                alt = otrack.track_sequence(kps, jf, *args[2:])
        for f in want:
            dq, dt = H.pose_close(want[f][0], alt[f][0])
            df = abs(float(alt[f][0].intrinsics.fy) - float(want[f][0].intrinsics.fy)) / abs(float(want[f][0].intrinsics.fy))
            band[f] = [max(band[f][0], BAND * dq), max(band[f][1], BAND * dt), max(band[f][2], BAND * df)]
    return want, band


def test_track_sequence_matches_oracle(ctx_small, scene):
    """SolveFrame chained over the clip (tracker.cc:133-192): GPU poses vs the oracle's."""
    from polychase_b200 import capi
    clip, kps, flows, NF = scene["clip"], scene["kps"], scene["flows"], scene["NF"]
    model = np.eye(4, dtype=F)
    ctx_small.mesh_set(scene["verts"], scene["tris"])
    opts = opnp.BundleOptions(loss_type=opnp.CAUCHY)
    want, band = _track_oracle_with_band(scene, opts)
    traj = {0: H.to_abi(cam_of(scene, 0))}
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
        assert om > 200
        assert abs(m - om) <= max(2, 0.002 * om)
        dq, dt = H.pose_close(ocam, H.from_abi(cam))
        assert dq <= max(RTOL, band[f][0]) and dt <= max(RTOL, band[f][1]), (f, dq, dt, band[f])
        assert abs(inl - oinl) < 5e-3
    # and the ground truth is recovered
    dq, dt = H.pose_close(cam_of(scene, NF - 1), H.from_abi(traj[NF - 1]))
    assert dq < 2e-3 and dt < 2e-3


@pytest.mark.parametrize("opt_f", [False, True])
def test_fused_analyze_track_chain_matches_oracle(ctx_small, scene, opt_f):
    """pc_analyze_track_begin: the forward sweep chained on the device behind the analyzer gives the
    oracle's TrackSequence poses (tracker.cc:133-213) and the same analyze rows as the plain pass."""
    from polychase_b200 import capi
    clip, kps, flows, NF = scene["clip"], scene["kps"], scene["flows"], scene["NF"]
    model = np.eye(4, dtype=F)
    opts = opnp.BundleOptions(loss_type=opnp.CAUCHY)
    start = cam_of(scene, 0)
    want, band = _track_oracle_with_band(scene, opts, opt_f)
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
        ctx_small.analyze_push(k, H.clip_rgb(clip, k, scene["conv"]))
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
        # with the focal length free, depth and focal length trade off along a nearly flat valley on this
        # almost planar scene and the chain feeds each pose to the next frame: the reference's own
        # summation-order noise (band) is what bounds the agreement there
        assert dq <= max(RTOL, band[f][0]) and dt <= max(RTOL, band[f][1]), (f, dq, dt, band[f])
        assert abs(r["inlier_ratio"] - oinl) < 5e-3
        assert abs(g.intrinsics.fy - ocam.intrinsics.fy) <= max(RTOL, band[f][2]) * abs(ocam.intrinsics.fy)


def test_fused_chain_not_enough_features(ctx_small, scene):
    """A frame whose rays all miss the mesh raises the reference's error (tracker.cc:160-166) at pop."""
    from polychase_b200 import capi
    clip, NF = scene["clip"], 3
    far = scene["verts"].copy()
    far[:, 0] += 1e4                                  # mesh out of view: every ray misses
    ctx_small.mesh_set(far, scene["tris"])
    ctx_small.analyze_begin(clip.width, clip.height, 0, NF, capi.default_gftt(max_corners=250))
    ctx_small.analyze_track_begin(np.eye(4, dtype=F), capi.default_bundle())
    ctx_small.analyze_track_seed(0, H.to_abi(cam_of(scene, 0)))
    ctx_small.analyze_push(0, H.clip_rgb(clip, 0, scene["conv"]))
    ctx_small.analyze_push(1, H.clip_rgb(clip, 1, scene["conv"]))
    assert ctx_small.analyze_pop()["tracked"] == 2
    with pytest.raises(capi.PcError) as e:
        ctx_small.analyze_pop()
    assert e.value.code == -7 and "Could not track to frame: 1" in str(e.value)
    ctx_small.analyze_end()
    ctx_small.mesh_set(scene["verts"], scene["tris"])


def test_track_not_enough_features(ctx_small, scene):
    from polychase_b200 import capi
    ctx_small.mesh_set(scene["verts"], scene["tris"])
    cam = H.to_abi(cam_of(scene, 0))
    with pytest.raises(capi.PcError) as e:
        ctx_small.track_frame([(cam, np.zeros((2, 2), F), np.array([0, 1], np.uint32), np.zeros((2, 2), F))],
                              np.eye(4, dtype=F), cam)
    assert e.value.code == -7


def ba_setup(scene, rng, opt_f=False, opt_pp=False, perturb=True, rot_deg=0.05, trans=0.004):
    kps, flows, NF = scene["kps"], scene["flows"], scene["NF"]
    first = scene.get("first", 0)
    model = np.eye(4, dtype=F)
    traj = [cam_of(scene, first + k) for k in range(NF)]
    if perturb:
        for k in range(1, NF - 1):
            traj[k] = H.perturb(traj[k], rng, rot_deg=rot_deg, trans=trans)
    edges = [oba.Edge(a - first, b - first, flows[(a, b)][0], flows[(a, b)][1]) for (a, b) in sorted(flows)
             if len(flows[(a, b)][0])]
    prob = oba.RefineProblem([kps[first + k] for k in range(NF)], edges, scene["verts"], scene["tris"], None, model,
                             opt_f, opt_pp, traj[0].intrinsics.bounds())
    abi_edges = [(e.src, e.tgt, e.src_kps_indices, e.tgt_kps) for e in edges]
    return model, traj, prob, abi_edges


def check_cost_and_normal_equations(ctx, scene, opt_f, opt_pp, loss, seed=3):
    from polychase_b200 import capi
    rng = np.random.default_rng(seed)
    model, traj, prob, abi_edges = ba_setup(scene, rng, opt_f, opt_pp)
    first = scene.get("first", 0)
    ctx.mesh_set(scene["verts"], scene["tris"])
    ctx.ba_load([scene["kps"][first + k] for k in range(scene["NF"])], abi_edges, model, opt_f, opt_pp)
    bo = capi.default_bundle(loss_type=loss)
    lo = opnp.Loss(loss, 1.0)
    atraj = [H.to_abi(c) for c in traj]
    with opnp.summation(np.float64):
        want_cost = prob.total_cost(traj, lo)
        A, g = prob.normal_equations(traj, lo)
    got_cost = ctx.ba_cost(atraj, bo)
    assert abs(got_cost - want_cost) <= RTOL * abs(want_cost), (got_cost, want_cost)
    # the primitive-id cache the cost pass leaves behind (refiner.cc:323-350) is the oracle's
    cache = ctx.ba_read_cache(int(prob.offs[-1]))
    want_cache = np.concatenate(prob.cache)
    assert (cache == want_cache).mean() > 0.998
    band, jtr = ctx.ba_normal_equations(atraj, bo)
    Ag = capi.band_to_dense(band)
    p = prob.p
    # per block: relative to that block's largest entry (a global max would hide the small blocks)
    nf = prob.nf
    worst = 0.0
    for i in range(nf):
        for j in range(max(0, i - 8), i + 1):
            blk_w = np.tril(A)[i * p:(i + 1) * p, j * p:(j + 1) * p]
            blk_g = Ag[i * p:(i + 1) * p, j * p:(j + 1) * p]
            s = np.abs(blk_w).max()
            if s == 0:
                assert not blk_g.any()
                continue
            worst = max(worst, float(np.abs(blk_g - blk_w).max() / s))
    assert worst <= RTOL, worst
    for i in range(nf):
        s = np.abs(g[i * p:(i + 1) * p]).max()
        if s > 0:
            assert np.abs(jtr[i * p:(i + 1) * p] - g[i * p:(i + 1) * p]).max() <= RTOL * max(s, 1e-3 * np.abs(g).max())
    # first / last frame are ground truth: their blocks are empty (refiner.cc:611-612)
    assert np.all(Ag[:p, :] == 0) and np.all(Ag[-p:, :] == 0)
    # K14 on its own (ComputeStep, lev_marq.h:826-841): the block-banded Cholesky solve of these equations against
    # a dense float64 solve of the same float32 matrix, at two dampings
    full = (Ag + np.tril(Ag, -1).T).astype(np.float64)
    dg = np.clip(np.diag(Ag), 1e-6, 1e32).astype(np.float64)
    for lam in (1e-5, 10.0):
        Ad = full.copy()
        np.fill_diagonal(Ad, dg * (1.0 + lam))
        want_step = -np.linalg.solve(Ad, jtr.astype(np.float64))
        got_step, got_norm = ctx.ba_solve_step(lam)
        # a float32 Cholesky is backward stable: the residual is small against |A| |x| + |b| whatever the conditioning
        # (with the intrinsics free the system is ill conditioned and the forward error alone says little)
        resid = Ad @ got_step.astype(np.float64) + jtr
        assert np.abs(resid).max() <= 2e-5 * (np.abs(Ad) @ np.abs(got_step) + np.abs(jtr)).max(), lam
        assert np.abs(got_step - want_step).max() <= (2e-3 if p == 6 else 5e-2) * np.abs(want_step).max(), lam
        assert abs(got_norm - np.linalg.norm(got_step.astype(np.float64))) <= 1e-5 * got_norm
    return prob


@pytest.mark.parametrize("opt_f,opt_pp,loss", [(False, False, 2), (True, True, 2), (False, False, 1)])
def test_ba_cost_and_normal_equations(ctx_small, scene, opt_f, opt_pp, loss):
    check_cost_and_normal_equations(ctx_small, scene, opt_f, opt_pp, loss)


def ba_oracle_with_band(scene, seed, opt_f, opt_pp, iters, seeds=(1, 2), **setup_kw):
    """Exact-sum oracle refine + the noise band of permuted float32 summation orders."""
    opts = opnp.BundleOptions(loss_type=opnp.CAUCHY, max_iterations=iters)
    model, traj, prob, abi_edges = ba_setup(scene, np.random.default_rng(seed), opt_f, opt_pp, **setup_kw)
    with opnp.summation(np.float64):
        want, wst = oba.refine_trajectory(prob, traj, opts)
    bq = bt = bc = 0.0
    for s in seeds:
        _, traj2, prob2, _ = ba_setup(scene, np.random.default_rng(seed), opt_f, opt_pp, **setup_kw)
        with opnp.summation(F, perm_seed=s):
            alt, ast = oba.refine_trajectory(prob2, traj2, opts)
        for k in range(len(traj)):
            dq, dt = H.pose_close(want[k], alt[k])
            bq, bt = max(bq, dq), max(bt, dt)
        bc = max(bc, abs(float(ast.cost) - float(wst.cost)) / abs(float(wst.cost)))
    return model, traj, abi_edges, want, wst, dict(q=BAND * bq, t=BAND * bt, cost=BAND * bc)


@pytest.mark.parametrize("opt_f,opt_pp", [(False, False), (True, False)])
def test_ba_solve_matches_oracle(ctx_small, scene, opt_f, opt_pp):
    from polychase_b200 import capi
    model, traj, abi_edges, want, wst, band = ba_oracle_with_band(scene, 5, opt_f, opt_pp, 30)
    ctx_small.mesh_set(scene["verts"], scene["tris"])
    ctx_small.ba_load([scene["kps"][k] for k in range(scene["NF"])], abi_edges, model, opt_f, opt_pp)
    seen = []
    got, gst = ctx_small.ba_solve([H.to_abi(c) for c in traj], capi.default_bundle(loss_type=2, max_iterations=30),
                                  callback=lambda s: seen.append(s.cost) or True)
    assert len(seen) >= 1
    assert abs(gst.initial_cost - wst.initial_cost) <= RTOL * abs(wst.initial_cost)
    assert gst.cost <= gst.initial_cost
    assert abs(gst.cost - wst.cost) <= max(RTOL, band["cost"]) * abs(wst.cost), (gst.cost, wst.cost, band)
    for k in range(len(traj)):
        dq, dt = H.pose_close(want[k], H.from_abi(got[k]))
        assert dq <= max(RTOL, band["q"]) and dt <= max(RTOL, band["t"]), (k, dq, dt, band)
        assert int(got[k].convention) == scene["conv"]
    # end frames untouched
    for k in (0, len(traj) - 1):
        assert list(got[k].q) == [float(v) for v in traj[k].pose.q]


def test_ba_empty_problem_returns_trajectory_unchanged(ctx_small, scene):
    """No keypoint inside the projected mesh bbox -> no edges, no residuals: RefineTrajectory leaves the
    trajectory as it is (refiner.cc:649-725 with an empty CachedDatabase) instead of failing a launch."""
    from polychase_b200 import capi
    NF = 4
    traj = [H.to_abi(cam_of(scene, k)) for k in range(NF)]
    ctx_small.mesh_set(scene["verts"], scene["tris"])
    ctx_small.ba_load([np.zeros((0, 2), F) for _ in range(NF)], [], np.eye(4, dtype=F), False, False)
    assert ctx_small.ba_cost(traj, capi.default_bundle(loss_type=2)) == 0.0
    got, st = ctx_small.ba_solve(traj, capi.default_bundle(loss_type=2, max_iterations=5))
    for k in range(NF):
        assert list(got[k].q) == list(traj[k].q) and list(got[k].t) == list(traj[k].t)


def test_ray_cast_nearest_hit_with_occluders(ctx_small):
    """Occlusion: the benchmark scene is one plane, so nothing there checks that the BVH returns the
    NEAREST hit.  Here the plane lies behind a box and a soup of 300 random triangles that overlap each other along
    the view rays, the model matrix is a general similarity, and a third of the primitives are masked out
    (ray_casting.cc:71-105: masked primitives are skipped, not opaque).  Against the brute-force oracle
    (oracle/raycast.py, every triangle tested): same hit flags, same primitive, and the hit distance is the minimum
    over all unmasked triangles."""
    w, h = 320, 240
    clip = synth.Clip(w, h, 4, seed=9)
    rng = np.random.default_rng(5)
    pv, pt = synth.plane_mesh(w, h, clip.s, quads=16)
    ext = np.abs(pv[:, :2]).max(0)
    # a box between the camera and the plane (the camera sits at z ~ -depth in object space and looks at +z ... or the
    # mirror image, depending on the convention: put occluders on both sides so that either way some are in front)
    def box(c, r):
        v = np.array([[x, y, z] for z in (-1, 1) for y in (-1, 1) for x in (-1, 1)], F) * r + c
        t = np.array([[0, 1, 3], [0, 3, 2], [4, 7, 5], [4, 6, 7], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6],
                      [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]], np.uint32)
        return v, t
    parts_v, parts_t, base = [pv.astype(F)], [pt.astype(np.uint32)], len(pv)
    for zc in (-0.8, 0.8):
        bv, bt = box(np.array([0.1 * ext[0], -0.05 * ext[1], zc], F), np.array([0.25 * ext[0], 0.2 * ext[1], 0.15], F))
        parts_v.append(bv); parts_t.append(bt + base); base += len(bv)
        c = np.stack([rng.uniform(-ext[0], ext[0], 150), rng.uniform(-ext[1], ext[1], 150), rng.uniform(0.2, 1.6, 150) * np.sign(zc)], 1)
        sv = (c[:, None, :] + rng.normal(0, 0.12 * ext[0], (150, 3, 3))).reshape(-1, 3).astype(F)
        parts_v.append(sv); parts_t.append(np.arange(450, dtype=np.uint32).reshape(150, 3) + base); base += len(sv)
    verts, tris = np.concatenate(parts_v), np.concatenate(parts_t)
    mask = np.zeros(((len(tris) + 31) // 32 + 3) // 4 * 4, np.uint32)          # padded to 4 words (geometry.h:63-65)
    for p in rng.choice(len(tris), len(tris) // 3, replace=False):
        mask[p // 32] |= np.uint32(1) << np.uint32(p % 32)
    ang = 0.3
    model = np.eye(4, dtype=F)
    model[:3, :3] = 1.1 * np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], F)
    model[:3, 3] = [0.05, -0.03, 0.1]
    pos = np.stack([rng.uniform(-10, w + 10, 6000), rng.uniform(-10, h + 10, 6000)], 1).astype(F)
    for conv in CONVENTIONS:
        cam = H.oracle_cam(clip, 1, conv)
        ctx_small.mesh_set(verts, tris, mask)
        hit, P, prim, uv, t = ctx_small.ray_cast(model, H.to_abi(cam), pos, True)
        o, d = oray.ray_object_space(model, cam.pose.Rt4x4(), cam.intrinsics, pos)
        eh, eP, eprim, euv, et = oray.ray_cast(verts, tris, mask, o, d, True)
        assert 0.3 < eh.mean()                                       # the camera sees the scene in this convention
        front = eprim[eh] >= len(pt)                                 # hits on the occluders, not on the plane behind them
        assert front.mean() > 0.2, front.mean()
        assert (hit == eh).mean() > 0.998
        both = hit & eh
        assert (prim[both] == eprim[both]).mean() > 0.995
        assert not np.any(mask[prim[both] // 32] >> (prim[both] % 32).astype(np.uint32) & 1)   # never a masked primitive
        # nearest: wherever the primitive differs (grazing rays), the distance still is the minimum to 1e-4
        assert np.percentile(np.abs(t[both] - et[both]) / np.maximum(np.abs(et[both]), 1e-6), 99.8) < 1e-4
        hit2, P2, prim2, uv2, t2 = ctx_small.ray_cast(model, H.to_abi(cam), pos, False)     # masks ignored
        eh2, _, eprim2, _, et2 = oray.ray_cast(verts, tris, mask, o, d, False)
        assert (hit2 == eh2).mean() > 0.998 and (prim2[hit2 & eh2] == eprim2[hit2 & eh2]).mean() > 0.995
        assert (et2[eh2 & eh] <= et[eh2 & eh] * (1 + 1e-6)).all()    # unmasking can only bring the hit nearer
    ctx_small.mesh_set(verts[:3], np.array([[0, 1, 2]], np.uint32))
