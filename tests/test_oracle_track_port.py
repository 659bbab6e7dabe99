"""Two independent restatements of the reference's Track path must agree: the numpy one
(oracle/pnp.py, raycast.py, track.py -- vectorised, brute-force ray casting) and the plain-C one
(oracle/track_port.c -- scalar loops, BVH), in both camera conventions.  CPU only."""
import numpy as np
import pytest

from oracle import geometry as G
from oracle import pnp as opnp
from oracle import raycast as oray
from oracle import synth
from oracle import track_port as tp
from tests import helpers as H

F = np.float32


@pytest.mark.parametrize("conv", [G.OPENCV, G.OPENGL])
def test_ray_cast_port_matches_bruteforce(conv):
    clip = synth.Clip(320, 240, 4, seed=4)
    verts, tris = H.bumpy_mesh(clip, quads=8, amp=0.03)
    mask = np.zeros(4, np.uint32)
    mask[0] = 0b1010_0000_0000
    cam = H.oracle_cam(clip, 2, conv)
    rng = np.random.default_rng(0)
    pos = np.stack([rng.uniform(-20, 340, 4000), rng.uniform(-20, 260, 4000)], 1).astype(F)
    model = np.eye(4, dtype=F)
    o, d = oray.ray_object_space(model, cam.pose.Rt4x4(), cam.intrinsics, pos)
    eh, eP, eprim, _, _ = oray.ray_cast(verts, tris, mask, o, d, True)
    hit, P, prim = tp.Mesh(verts, tris, mask).ray_cast(model, cam, pos, True)
    assert eh.mean() > 0.5
    assert (hit == eh).mean() > 0.999
    both = hit & eh
    assert (prim[both] == eprim[both]).mean() > 0.998
    assert np.percentile(np.abs(P[both] - eP[both]).max(1), 99.9) < 1e-5 * clip.depth


@pytest.mark.parametrize("conv", [G.OPENCV, G.OPENGL])
@pytest.mark.parametrize("loss,opt_f,opt_pp", [(0, False, False), (1, False, False), (2, False, False), (2, True, True)])
def test_solve_pnp_port_matches_numpy_restatement(conv, loss, opt_f, opt_pp):
    clip = synth.Clip(320, 240, 6, seed=4)
    gt = H.oracle_cam(clip, 5, conv)
    rng = np.random.default_rng(7)
    n = 800
    X = np.stack([rng.uniform(-1.5, 1.5, n), rng.uniform(-1.0, 1.0, n), rng.uniform(-0.05, 0.05, n)], 1).astype(F)
    x = gt.intrinsics.project(gt.pose.apply(X)) + rng.normal(0, 0.3, (n, 2)).astype(F)
    x[:20] += rng.normal(0, 25, (20, 2)).astype(F)
    init = H.perturb(gt, rng)
    opts = opnp.BundleOptions(loss_type=loss)
    a, ast, ainl = opnp.solve_pnp_iterative(X, x, None, init, opts, 12.0, opt_f, opt_pp)
    b, bst, binl = tp.solve_pnp(X, x, None, init, opts, 12.0, opt_f, opt_pp)
    dq, dt = H.pose_close(a, b)
    tol = 1e-4 if not opt_f else 1e-3          # free focal length: flat valley, the LM paths part within rounding
    assert dq < tol and dt < tol, (dq, dt)
    assert abs(float(ast.initial_cost) - float(bst.initial_cost)) <= 1e-5 * abs(float(ast.initial_cost))
    assert abs(float(ast.cost) - float(bst.cost)) <= 1e-3 * abs(float(ast.cost))
    assert abs(float(ainl) - float(binl)) < 2e-3
    assert abs(float(a.intrinsics.fy) - float(b.intrinsics.fy)) <= tol * abs(float(a.intrinsics.fy))


def test_track_frame_port_recovers_ground_truth():
    clip = synth.Clip(320, 240, 6, seed=4)
    verts, tris = synth.plane_mesh(320, 240, clip.s, quads=4)
    mesh = tp.Mesh(verts, tris)
    model = np.eye(4, dtype=F)
    rng = np.random.default_rng(3)
    srcs = []
    tgt_cam = H.oracle_cam(clip, 4)
    for s in (3, 2, 0):
        cam = H.oracle_cam(clip, s)
        kps = np.stack([rng.uniform(20, 300, 300), rng.uniform(20, 220, 300)], 1).astype(F)
        o, d = oray.ray_object_space(model, cam.pose.Rt4x4(), cam.intrinsics, kps)
        hit, P, _, _, _ = oray.ray_cast(verts, tris, None, o, d, True)
        tgt = tgt_cam.intrinsics.project(tgt_cam.pose.apply(P)) + rng.normal(0, 0.2, (300, 2)).astype(F)
        idx = np.nonzero(hit)[0].astype(np.uint32)
        srcs.append((cam, kps, idx, tgt[hit]))
    out = tp.track_frame(mesh, model, srcs, H.oracle_cam(clip, 3), opnp.BundleOptions(loss_type=opnp.CAUCHY))
    assert out is not None
    cam, st, inl, m = out
    assert m > 800 and inl > 0.99
    dq, dt = H.pose_close(tgt_cam, cam)
    assert dq < 5e-4 and dt < 5e-4
    far = verts.copy()
    far[:, 0] += 1e4
    assert tp.track_frame(tp.Mesh(far, tris), model, srcs, H.oracle_cam(clip, 3), opnp.BundleOptions()) is None
