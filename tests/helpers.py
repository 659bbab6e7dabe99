"""Shared test helpers: synthetic scenes, conversions between the oracle's and the C ABI's
camera records."""
from __future__ import annotations

import numpy as np

from oracle import geometry as G
from oracle import synth

F = np.float32


# The addon's camera (blender_addon/core.py:348-357): OpenGL convention (looks down -Z,
# cpp/pnp/types.h:13-16) with fx, fy negated, on images whose rows run bottom-up
# (blender_addon/operators/analysis.py:220-233).  The same physical camera as the clip's OpenCV one:
# camera axes turned by pi about x (y and z flip), image rows flipped (y -> H-1-y), so that
#   fx_gl * Xg.x / Xg.z + cx = u   and   fy_gl * Xg.y / Xg.z + (H-1-cy) = H-1-v.
GL_FLIP = np.diag([1.0, -1.0, -1.0])


def oracle_cam(clip: synth.Clip, k: int, convention=G.OPENCV) -> G.CameraState:
    K = clip.K
    i = k - clip.first_frame
    if convention == G.OPENCV:
        it = G.Intrinsics(K["fx"], K["fy"], K["cx"], K["cy"], 1.0, clip.width, clip.height, convention).f32()
        return G.CameraState(it, G.Pose(G.quat_from_matrix(clip.R[i]).astype(F), clip.t[i].astype(F)))
    it = G.Intrinsics(-K["fx"], -K["fy"], K["cx"], (clip.height - 1) - K["cy"], 1.0, clip.width, clip.height,
                      G.OPENGL).f32()
    return G.CameraState(it, G.Pose(G.quat_from_matrix(GL_FLIP @ clip.R[i]).astype(F),
                                    (GL_FLIP @ clip.t[i]).astype(F)))


def clip_rgb(clip: synth.Clip, k: int, convention=G.OPENCV) -> np.ndarray:
    """Frame k as the given camera convention sees it (OpenGL: rows bottom-up, like Blender's buffers)."""
    rgb = clip.rgb(k)
    return rgb if convention == G.OPENCV else np.ascontiguousarray(rgb[::-1])


def to_abi(cam: G.CameraState):
    from polychase_b200 import capi
    it = cam.intrinsics
    cs = capi.CameraState(float(it.fx), float(it.fy), float(it.cx), float(it.cy), float(it.aspect_ratio),
                          float(it.width), float(it.height), float(it.convention))
    cs.q[:] = [float(v) for v in cam.pose.q]
    cs.t[:] = [float(v) for v in cam.pose.t]
    cs.filled = 1.0
    return cs


def from_abi(cs) -> G.CameraState:
    it = G.Intrinsics(F(cs.fx), F(cs.fy), F(cs.cx), F(cs.cy), F(cs.aspect_ratio), F(cs.width), F(cs.height),
                      int(cs.convention))
    return G.CameraState(it, G.Pose(np.array(list(cs.q), F), np.array(list(cs.t), F)))


def perturb(cam: G.CameraState, rng, rot_deg=0.2, trans=0.02) -> G.CameraState:
    out = cam.copy()
    w = rng.normal(0, np.deg2rad(rot_deg), 3).astype(F)
    out.pose.q = G.quat_step_post(cam.pose.q, w)
    out.pose.t = (cam.pose.t + rng.normal(0, trans, 3)).astype(F)
    return out


def pose_close(a: G.CameraState, b: G.CameraState, rtol=1e-4):
    qa, qb = np.asarray(a.pose.q, np.float64), np.asarray(b.pose.q, np.float64)
    if np.dot(qa, qb) < 0:
        qb = -qb
    dq = np.abs(qa - qb).max()
    scale = max(np.abs(np.asarray(a.pose.t, np.float64)).max(), 1e-6)
    dt = np.abs(np.asarray(a.pose.t, np.float64) - np.asarray(b.pose.t, np.float64)).max() / scale
    return dq, dt


def bumpy_mesh(clip: synth.Clip, quads=8, amp=0.05, seed=0):
    """A small non-planar mesh under the clip's view (so occlusion / several candidates exist)."""
    verts, tris = synth.plane_mesh(clip.width, clip.height, clip.s, quads=quads)
    rng = np.random.default_rng(seed)
    verts = verts.copy()
    verts[:, 2] = rng.uniform(-amp, amp, len(verts)).astype(F)
    return verts, tris
