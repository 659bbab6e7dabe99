"""ctypes front-end of oracle/track_port.c (plain-C restatement of the Track sweep) -- TEST INFRASTRUCTURE
and the Track half of bench.py's CPU arm.  See that file's header for what it restates."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import restate
from .geometry import CameraState, Intrinsics, Pose
from .pnp import BundleOptions, BundleStats

F = np.float32


class Cam(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("aspect", C.c_float),
                ("width", C.c_float), ("height", C.c_float), ("convention", C.c_float), ("q", C.c_float * 4),
                ("t", C.c_float * 3), ("filled", C.c_float)]


class Opts(C.Structure):
    _fields_ = [("max_iterations", C.c_uint64), ("loss_type", C.c_int), ("loss_scale", C.c_float),
                ("gradient_tol", C.c_float), ("step_tol", C.c_float), ("initial_lambda", C.c_float),
                ("min_lambda", C.c_float), ("max_lambda", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("iterations", C.c_uint64), ("initial_cost", C.c_float), ("cost", C.c_float), ("lambda_", C.c_float),
                ("invalid_steps", C.c_uint64), ("step_norm", C.c_float), ("grad_norm", C.c_float)]


_ready = False


def _lib():
    global _ready
    lib = restate.lib()
    if not _ready:
        lib.orc_bvh_build.restype = C.c_void_p
        lib.orc_bvh_build.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.orc_bvh_free.argtypes = [C.c_void_p]
        lib.orc_ray_cast.restype = C.c_int
        lib.orc_ray_cast.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Cam), C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_solve_pnp.restype = C.c_int
        lib.orc_solve_pnp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Opts), C.c_float, C.c_int,
                                      C.c_int, C.POINTER(Cam), C.POINTER(Stats), C.POINTER(C.c_float)]
        lib.orc_track_frame.restype = C.c_int
        lib.orc_track_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Cam),
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_int), C.POINTER(Cam), C.POINTER(Opts), C.c_int, C.c_int,
                                        C.POINTER(Cam), C.POINTER(Stats), C.POINTER(C.c_float)]
        _ready = True
    return lib


def to_cam(cs: CameraState) -> Cam:
    it = cs.intrinsics
    c = Cam(float(it.fx), float(it.fy), float(it.cx), float(it.cy), float(it.aspect_ratio), float(it.width),
            float(it.height), float(it.convention))
    c.q[:] = [float(v) for v in cs.pose.q]
    c.t[:] = [float(v) for v in cs.pose.t]
    c.filled = 1.0
    return c


def from_cam(c: Cam) -> CameraState:
    it = Intrinsics(F(c.fx), F(c.fy), F(c.cx), F(c.cy), F(c.aspect), F(c.width), F(c.height), int(c.convention))
    return CameraState(it, Pose(np.array(list(c.q), F), np.array(list(c.t), F)))


def to_opts(o: BundleOptions) -> Opts:
    return Opts(int(o.max_iterations), int(o.loss_type), float(o.loss_scale), float(o.gradient_tol), float(o.step_tol),
                float(o.initial_lambda), float(o.min_lambda), float(o.max_lambda))


def _stats(s: Stats) -> BundleStats:
    return BundleStats(int(s.iterations), F(s.initial_cost), F(s.cost), F(s.lambda_), int(s.invalid_steps),
                       F(s.step_norm), F(s.grad_norm))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Mesh:
    """BVH over a triangle mesh (stand-in for the reference's AcceleratedMesh, ray_casting.cc:21-63)."""

    def __init__(self, verts, tris, mask_bits=None):
        self.verts = np.ascontiguousarray(verts, F).reshape(-1, 3)
        self.tris = np.ascontiguousarray(tris, np.uint32).reshape(-1, 3)
        self.mask = None if mask_bits is None or len(mask_bits) == 0 else np.ascontiguousarray(mask_bits, np.uint32)
        self.h = _lib().orc_bvh_build(_p(self.verts), len(self.verts), _p(self.tris), len(self.tris))

    def __del__(self):
        try:
            if self.h:
                _lib().orc_bvh_free(self.h)
                self.h = None
        except Exception:
            pass

    def ray_cast(self, model, cam: CameraState, pos, check_mask=True):
        model = np.ascontiguousarray(model, F).reshape(16)
        pos = np.ascontiguousarray(pos, F).reshape(-1, 2)
        n = len(pos)
        hit = np.zeros(n, np.uint8)
        P = np.zeros((n, 3), F)
        prim = np.zeros(n, np.uint32)
        c = to_cam(cam)
        rc = _lib().orc_ray_cast(self.h, _p(model), C.byref(c), _p(pos), n, _p(self.mask), int(check_mask), _p(hit), _p(P),
                                 _p(prim))
        assert rc == 0
        return hit.astype(bool), P, prim


def solve_pnp(X, x, weights, cam: CameraState, opts: BundleOptions, max_inlier_error=12.0, opt_f=False, opt_pp=False):
    X = np.ascontiguousarray(X, F).reshape(-1, 3)
    x = np.ascontiguousarray(x, F).reshape(-1, 2)
    w = None if weights is None else np.ascontiguousarray(weights, F)
    c = to_cam(cam)
    st = Stats()
    inl = C.c_float()
    o = to_opts(opts)
    rc = _lib().orc_solve_pnp(_p(X), _p(x), _p(w), len(X), C.byref(o), float(max_inlier_error), int(opt_f), int(opt_pp),
                              C.byref(c), C.byref(st), C.byref(inl))
    assert rc == 0
    return from_cam(c), _stats(st), F(inl.value)


def track_frame(mesh: Mesh, model, sources, init: CameraState, opts: BundleOptions, opt_f=False, opt_pp=False):
    """sources: list of (CameraState, keypoints (nk,2), src_idx (rows,), tgt (rows,2)).
    Returns (CameraState, BundleStats, inlier_ratio, matches) or None when fewer than 3 rays hit."""
    model = np.ascontiguousarray(model, F).reshape(16)
    n = len(sources)
    cams = (Cam * max(n, 1))()
    kps_p = (C.c_void_p * max(n, 1))()
    idx_p = (C.c_void_p * max(n, 1))()
    tgt_p = (C.c_void_p * max(n, 1))()
    rows = (C.c_int * max(n, 1))()
    keep = []
    for i, (cam, kps, idx, tgt) in enumerate(sources):
        kps = np.ascontiguousarray(kps, F).reshape(-1, 2)
        idx = np.ascontiguousarray(idx, np.uint32)
        tgt = np.ascontiguousarray(tgt, F).reshape(-1, 2)
        keep += [kps, idx, tgt]
        cams[i] = to_cam(cam)
        kps_p[i], idx_p[i], tgt_p[i], rows[i] = kps.ctypes.data, idx.ctypes.data, tgt.ctypes.data, len(idx)
    out = Cam()
    st = Stats()
    inl = C.c_float()
    o = to_opts(opts)
    ci = to_cam(init)
    m = _lib().orc_track_frame(mesh.h, _p(mesh.mask), _p(model), n, cams, kps_p, idx_p, tgt_p, rows, C.byref(ci), C.byref(o),
                               int(opt_f), int(opt_pp), C.byref(out), C.byref(st), C.byref(inl))
    if m < 3:
        return None
    return from_cam(out), _stats(st), F(inl.value), m
