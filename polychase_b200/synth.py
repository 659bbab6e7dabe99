"""Deterministic synthetic clip generator (numpy only, no cv2): the INPUT side of bench.py, the scripts and
the tests (textures, camera paths, homographies, the plane mesh).  Not part of the hot path and not
the checker; the device-side renderer that pairs with it is pc_synth_render (csrc/kernels/synth.cu).

Scene: a textured planar mesh at z=0 seen by a pinhole
camera on a smooth path.  Frame k is the homography warp of one texture, sampled
bilinearly with REFLECT_101 borders, stored as RGB u8 with R=G=B.  Poses are the
object->camera (model-view) rigid transforms the reference's CameraState.pose holds
(cpp/pnp/types.h:195-198, Appendix C of SURVEY.md), OpenCV convention
(camera looks down +Z, cpp/pnp/types.h:13-16).
"""
from __future__ import annotations

import numpy as np


def make_texture(width: int, height: int, seed: int = 0) -> np.ndarray:
    """Band-limited noise texture, u8 (height, width)."""
    rng = np.random.default_rng(seed)
    lw, lh = (width + 3) // 4 + 2, (height + 3) // 4 + 2
    low = rng.integers(0, 256, size=(lh, lw)).astype(np.float32)
    up = np.kron(low, np.ones((4, 4), np.float32))[: height + 8, : width + 8]
    k = np.array([1, 4, 6, 4, 1], np.float32) / 16.0
    for _ in range(3):
        up = (
            k[0] * up[:, :-4] + k[1] * up[:, 1:-3] + k[2] * up[:, 2:-2]
            + k[3] * up[:, 3:-1] + k[4] * up[:, 4:]
        )
        up = np.pad(up, ((0, 0), (2, 2)), mode="reflect")
        up = (
            k[0] * up[:-4] + k[1] * up[1:-3] + k[2] * up[2:-2]
            + k[3] * up[3:-1] + k[4] * up[4:]
        )
        up = np.pad(up, ((2, 2), (0, 0)), mode="reflect")
    up = up[4 : 4 + height, 4 : 4 + width]
    # stretch contrast back to the full range
    lo, hi = np.percentile(up, [1, 99])
    up = (up - lo) / max(hi - lo, 1e-6) * 255.0
    return np.clip(np.rint(up), 0, 255).astype(np.uint8)


def rot_xyz(rx: float, ry: float, rz: float) -> np.ndarray:
    cx, sx = np.cos(rx), np.sin(rx)
    cy, sy = np.cos(ry), np.sin(ry)
    cz, sz = np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def survey_speed(width: int) -> float:
    """Time scale of `camera_path` that keeps the largest image motion over 8 frames at about 20 px
    for a `width`-pixel frame -- SURVEY.md section 8d: "max inter-frame flow at skip 8 <~ 20 px,
    inside the 4-level pyramid's capture range".  At speed 1 the +-2 degree yaw/pitch sweeps move the
    image by up to 89 px per 8 frames at 4K (f = 1.2 W = 4608 px: one degree is 80 px), which is
    outside what a 4-level pyramid with 10x10 windows can follow: 54 % of the skip-4 and 96 % of the
    skip-8 tracks then end more than 3 px from the true position (profiles/r2_lk_track_error_by_skip.txt)."""
    return min(1.0, 20.0 / (89.3 * width / 3840.0))


def camera_path(num_frames: int, depth: float = 4.0, first: int = 0, speed: float = 1.0):
    """Smooth object->camera poses (R, t), float64.  yaw/pitch +-2 deg, translation
    <= 2% of depth per 8 frames.  `speed` scales time (see survey_speed)."""
    Rs, ts = [], []
    for k in range(first, first + num_frames):
        a = 2.0 * np.pi * k * speed / 97.0
        b = 2.0 * np.pi * k * speed / 61.0
        R = rot_xyz(np.deg2rad(2.0) * np.sin(b), np.deg2rad(2.0) * np.sin(a),
                    np.deg2rad(1.0) * np.sin(0.5 * a))
        t = np.array([
            0.02 * depth * 1.5 * np.sin(a),
            0.02 * depth * 1.0 * np.sin(b + 0.7),
            depth * (1.0 + 0.03 * np.sin(0.7 * a)),
        ])
        Rs.append(R)
        ts.append(t)
    return np.stack(Rs), np.stack(ts)


def intrinsics(width: int, height: int):
    """fx=fy=1.2*W, cx=W/2, cy=H/2, OpenCV convention."""
    f = 1.2 * width
    return dict(fx=f, fy=f, cx=width / 2.0, cy=height / 2.0, aspect_ratio=1.0,
                width=float(width), height=float(height), convention=1)


def plane_scale(width: int, depth: float = 4.0) -> float:
    """World units per texture pixel so the plane fills the view at `depth`."""
    return depth / (1.2 * width) * 1.15


def homography(K: dict, R: np.ndarray, t: np.ndarray, width: int, height: int,
               s: float) -> np.ndarray:
    """Texture pixel (u, v, 1) -> image pixel, for the plane z=0 with
    X = (u - W/2) * s, Y = (v - H/2) * s."""
    Km = np.array([[K["fx"], 0, K["cx"]], [0, K["fy"], K["cy"]], [0, 0, 1.0]])
    A = np.array([[s, 0, -s * width / 2.0], [0, s, -s * height / 2.0], [0, 0, 1.0]])
    P = np.stack([R[:, 0], R[:, 1], t], axis=1)
    return Km @ P @ A


def reflect101(idx: np.ndarray, n: int) -> np.ndarray:
    if n == 1:
        return np.zeros_like(idx)
    period = 2 * (n - 1)
    m = np.mod(idx, period)
    return np.where(m >= n, period - m, m)


def warp_frame(tex: np.ndarray, H: np.ndarray) -> np.ndarray:
    """Inverse-warp `tex` by homography H (texture->image); bilinear, REFLECT_101."""
    h, w = tex.shape
    Hi = np.linalg.inv(H)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    den = Hi[2, 0] * xs + Hi[2, 1] * ys + Hi[2, 2]
    u = (Hi[0, 0] * xs + Hi[0, 1] * ys + Hi[0, 2]) / den
    v = (Hi[1, 0] * xs + Hi[1, 1] * ys + Hi[1, 2]) / den
    u0 = np.floor(u).astype(np.int64)
    v0 = np.floor(v).astype(np.int64)
    fu = (u - u0).astype(np.float32)
    fv = (v - v0).astype(np.float32)
    x0, x1 = reflect101(u0, w), reflect101(u0 + 1, w)
    y0, y1 = reflect101(v0, h), reflect101(v0 + 1, h)
    t = tex.astype(np.float32)
    top = t[y0, x0] * (1 - fu) + t[y0, x1] * fu
    bot = t[y1, x0] * (1 - fu) + t[y1, x1] * fu
    out = top * (1 - fv) + bot * fv
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def plane_mesh(width: int, height: int, s: float, quads: int = 64, margin: float = 1.6):
    """z=0 grid mesh, quads x quads quads = 2*quads^2 triangles, spanning `margin` x
    the texture extent so every view ray hits it.  float32 verts, uint32 tris."""
    hx, hy = 0.5 * width * s * margin, 0.5 * height * s * margin
    gx = np.linspace(-hx, hx, quads + 1)
    gy = np.linspace(-hy, hy, quads + 1)
    X, Y = np.meshgrid(gx, gy)
    verts = np.stack([X.ravel(), Y.ravel(), np.zeros(X.size)], axis=1).astype(np.float32)
    tris = []
    n = quads + 1
    for j in range(quads):
        for i in range(quads):
            a = j * n + i
            tris.append((a, a + 1, a + n))
            tris.append((a + 1, a + n + 1, a + n))
    return verts, np.asarray(tris, np.uint32)


class Clip:
    """A deterministic synthetic clip: frames on demand, ground-truth cameras, mesh."""

    def __init__(self, width: int, height: int, num_frames: int, seed: int = 0,
                 first_frame: int = 0, depth: float = 4.0, speed: float = 1.0):
        self.width, self.height, self.num_frames = width, height, num_frames
        self.first_frame = first_frame
        self.depth = depth
        self.tex = make_texture(width, height, seed)
        self.K = intrinsics(width, height)
        self.s = plane_scale(width, depth)
        self.speed = speed
        self.R, self.t = camera_path(num_frames, depth, first_frame, speed)
        self.verts, self.tris = plane_mesh(width, height, self.s)

    def homography(self, k: int) -> np.ndarray:
        i = k - self.first_frame
        return homography(self.K, self.R[i], self.t[i], self.width, self.height, self.s)

    def gray(self, k: int) -> np.ndarray:
        return warp_frame(self.tex, self.homography(k))

    def rgb(self, k: int) -> np.ndarray:
        g = self.gray(k)
        return np.ascontiguousarray(np.repeat(g[:, :, None], 3, axis=2))
