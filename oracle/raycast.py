"""Brute-force nearest-hit ray casting -- TEST INFRASTRUCTURE.

PARITY UNPINNED against Embree itself (absent); exact against the analytic plane / brute force.

Stands in for Embree's rtcIntersect1 (third party, Embree 4.3.1, not in /root/reference) as
AcceleratedMesh::RayCast uses it (/root/reference/cpp/ray_casting.cc:65-121): nearest hit
with tnear = 0, a hit on a masked triangle is a miss (it does not continue behind it,
ray_casting.cc:106-108), position = (1-u-v) p1 + u p2 + v p3 (geometry.h:17-19).
Intersection test: Moller-Trumbore as in /root/reference/cpp/ray_casting.h:125-179."""
from __future__ import annotations

import numpy as np

F = np.float32


def ray_object_space(model, view, intr, pos):
    """GetRayObjectSpace, ray_casting.h:53-63 (4x4 inverse formed in float64, cast to f32)."""
    mat = np.linalg.inv((np.asarray(view, np.float64) @ np.asarray(model, np.float64))).astype(F)
    origin = mat[:3, 3].copy()
    d = (intr.unproject(np.asarray(pos, F)) @ mat[:3, :3].T).astype(F)
    return origin, d


def moller_trumbore(origin, dirs, p1, p2, p3):
    """Vectorised over rays (N,3) x triangles (M,3).  Returns t (N,M) with inf for misses, u, v."""
    eps = F(1e-10)
    e1 = (p2 - p1).astype(F)[None]                     # (1,M,3)
    e2 = (p3 - p1).astype(F)[None]
    d = dirs[:, None, :].astype(F)                     # (N,1,3)
    rxe2 = np.cross(d, e2).astype(F)
    det = np.sum(e1 * rxe2, -1, dtype=F)
    ok = ~((det > -eps) & (det < eps))
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = (F(1.0) / det).astype(F)
        s = (origin[None, None, :] - p1[None]).astype(F) if origin.ndim == 1 else (origin[:, None, :] - p1[None]).astype(F)
        u = (inv * np.sum(s * rxe2, -1, dtype=F)).astype(F)
        ok &= ~((u < 0) | (u > 1))
        sxe1 = np.cross(s, e1).astype(F)
        v = (inv * np.sum(d * sxe1, -1, dtype=F)).astype(F)
        ok &= ~((v < 0) | (u + v > 1))
        t = (inv * np.sum(e2 * sxe1, -1, dtype=F)).astype(F)
        ok &= ~(t < 0)
    t = np.where(ok, t, np.inf).astype(F)
    return t, u, v


def ray_cast(verts, tris, mask_bits, origin, dirs, check_mask=True, chunk=2048):
    """Nearest hit per ray.  Returns hit (N,) bool, pos (N,3) object space, prim (N,) u32, uv, t."""
    verts = np.asarray(verts, F)
    tris = np.asarray(tris, np.int64)
    p1, p2, p3 = verts[tris[:, 0]], verts[tris[:, 1]], verts[tris[:, 2]]
    n = len(dirs)
    hit = np.zeros(n, bool)
    pos = np.zeros((n, 3), F)
    prim = np.full(n, 0xFFFFFFFF, np.uint32)
    uv = np.zeros((n, 2), F)
    tt = np.zeros(n, F)
    for a in range(0, n, chunk):
        d = np.asarray(dirs[a:a + chunk], F)
        o = origin if origin.ndim == 1 else np.asarray(origin[a:a + chunk], F)
        t, u, v = moller_trumbore(o, d, p1, p2, p3)
        j = np.argmin(t, axis=1)
        r = np.arange(len(d))
        tb = t[r, j]
        h = np.isfinite(tb)
        if check_mask and mask_bits is not None and len(mask_bits):
            mb = np.asarray(mask_bits, np.uint32)
            masked = (mb[j // 32] >> (j % 32).astype(np.uint32)) & 1
            h &= masked == 0
        ub, vb = u[r, j], v[r, j]
        P = ((F(1.0) - ub - vb)[:, None] * p1[j] + ub[:, None] * p2[j] + vb[:, None] * p3[j]).astype(F)
        hit[a:a + chunk] = h
        pos[a:a + chunk] = np.where(h[:, None], P, 0)
        prim[a:a + chunk] = np.where(h, j, 0xFFFFFFFF).astype(np.uint32)
        uv[a:a + chunk] = np.stack([ub, vb], 1)
        tt[a:a + chunk] = np.where(h, tb, 0)
    return hit, pos, prim, uv, tt
