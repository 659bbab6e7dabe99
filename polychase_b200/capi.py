"""ctypes binding of libpolychase_b200.so (include/polychase_b200.h).

This is the Python face of the C ABI used by tests, bench.py and the smoke entry; the
reference-shaped surface (`polychase_core` classes) lives in polychase_b200/core.py and
the pybind11 module.  No CPU fallback: `load()` raises if the library is missing, and
`Context()` raises if no sm_100 device is usable.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpolychase_b200.so")
_lib = None

PC_MEM_HOST, PC_MEM_DEVICE, PC_MEM_HOST_PINNED = 0, 1, 2


class PcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"polychase_b200 error {code}: {msg}")
        self.code = code


class Limits(C.Structure):
    _fields_ = [("device", C.c_int), ("max_width", C.c_int), ("max_height", C.c_int),
                ("max_features", C.c_int), ("ring_frames", C.c_int), ("pipeline_depth", C.c_int)]


class GfttOpts(C.Structure):
    _fields_ = [("quality_level", C.c_double), ("min_distance", C.c_double), ("block_size", C.c_int),
                ("gradient_size", C.c_int), ("max_corners", C.c_int), ("use_harris", C.c_int),
                ("harris_k", C.c_double), ("grid_rows", C.c_int), ("grid_cols", C.c_int)]


class FlowOpts(C.Structure):
    _fields_ = [("window_size", C.c_int), ("max_level", C.c_int), ("term_max_iters", C.c_int),
                ("term_epsilon", C.c_double), ("min_eigen_threshold", C.c_double)]


class VideoInfo(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("first_frame", C.c_int32),
                ("num_frames", C.c_uint32)]


class CameraState(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("aspect_ratio", C.c_float), ("width", C.c_float), ("height", C.c_float),
                ("convention", C.c_float), ("q", C.c_float * 4), ("t", C.c_float * 3),
                ("filled", C.c_float)]


class BundleOpts(C.Structure):
    _fields_ = [("max_iterations", C.c_uint64), ("max_allowed_parallelism", C.c_uint64),
                ("loss_type", C.c_int), ("loss_scale", C.c_float), ("gradient_tol", C.c_float),
                ("step_tol", C.c_float), ("initial_lambda", C.c_float), ("min_lambda", C.c_float),
                ("max_lambda", C.c_float), ("verbose", C.c_int)]


class BundleStats(C.Structure):
    _fields_ = [("iterations", C.c_uint64), ("initial_cost", C.c_float), ("cost", C.c_float),
                ("lambda_", C.c_float), ("invalid_steps", C.c_uint64), ("step_norm", C.c_float),
                ("grad_norm", C.c_float)]


class PairRows(C.Structure):
    _fields_ = [("image_id_from", C.c_int32), ("image_id_to", C.c_int32), ("rows", C.c_int32),
                ("src_kps_indices", C.POINTER(C.c_uint32)), ("tgt_kps", C.POINTER(C.c_float)),
                ("flow_errors", C.POINTER(C.c_float))]


class FrameResult(C.Structure):
    _fields_ = [("frame_id", C.c_int32), ("num_keypoints", C.c_int32), ("keypoints", C.POINTER(C.c_float)),
                ("num_pairs", C.c_int32), ("pairs", PairRows * 8),
                ("tracked", C.c_int32), ("num_matches", C.c_int32), ("inlier_ratio", C.c_float),
                ("camera", CameraState), ("stats", BundleStats)]


class KernelTimes(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("gray_pyr_ms", "min_eig_ms", "select_ms", "lk_ms", "compact_ms",
                                           "raycast_ms", "pnp_ms", "ba_ms", "lk_tmpl_ms")] + \
               [(n, C.c_uint64) for n in ("gray_pyr_n", "min_eig_n", "select_n", "lk_n", "compact_n",
                                           "raycast_n", "pnp_n", "ba_n", "lk_tmpl_n")]


class MatchSource(C.Structure):
    _fields_ = [("camera", CameraState), ("keypoints", C.c_void_p), ("nk", C.c_int32),
                ("src_kps_indices", C.c_void_p), ("tgt_kps", C.c_void_p), ("rows", C.c_int32)]


class BAEdge(C.Structure):
    _fields_ = [("src_frame_idx", C.c_int32), ("tgt_frame_idx", C.c_int32), ("first_row", C.c_int32),
                ("rows", C.c_int32)]


class BAProblem(C.Structure):
    _fields_ = [("num_frames", C.c_int32), ("kp_offsets", C.c_void_p), ("keypoints", C.c_void_p),
                ("num_edges", C.c_int32), ("edges", C.c_void_p), ("src_kps_indices", C.c_void_p),
                ("tgt_kps", C.c_void_p), ("model", C.c_float * 16), ("optimize_focal_length", C.c_int),
                ("optimize_principal_point", C.c_int)]


BA_ITER_CB = C.CFUNCTYPE(C.c_int, C.POINTER(BundleStats), C.c_void_p)

# name -> (restype, argtypes); every symbol include/polychase_b200.h declares
SIGNATURES = {
    "pc_default_gftt_opts": (None, [C.POINTER(GfttOpts)]),
    "pc_default_flow_opts": (None, [C.POINTER(FlowOpts)]),
    "pc_default_bundle_opts": (None, [C.POINTER(BundleOpts)]),
    "pc_version": (C.c_char_p, []),
    "pc_create": (C.c_int, [C.POINTER(Limits), C.POINTER(C.c_void_p)]),
    "pc_destroy": (None, [C.c_void_p]),
    "pc_last_error": (C.c_char_p, [C.c_void_p]),
    "pc_synchronize": (C.c_int, [C.c_void_p]),
    "pc_kernel_launches": (C.c_uint64, [C.c_void_p]),
    "pc_frame_upload_rgb8": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.POINTER(FlowOpts)]),
    "pc_frame_from_device_rgb8": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.POINTER(FlowOpts)]),
    "pc_frame_upload_gray8": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.POINTER(FlowOpts)]),
    "pc_frame_release": (C.c_int, [C.c_void_p, C.c_int32]),
    "pc_frame_num_levels": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int)]),
    "pc_frame_read_level": (C.c_int, [C.c_void_p, C.c_int32, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pc_detect": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(GfttOpts), C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "pc_min_eig_map": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t]),
    "pc_set_keypoints": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int]),
    "pc_lk_pair": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(FlowOpts), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "pc_lk_raw": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(FlowOpts), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "pc_analyze_begin": (C.c_int, [C.c_void_p, C.POINTER(VideoInfo), C.POINTER(GfttOpts), C.POINTER(FlowOpts)]),
    "pc_analyze_push_frame": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_int]),
    "pc_analyze_preset_keypoints": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int]),
    "pc_analyze_pop": (C.c_int, [C.c_void_p, C.POINTER(FrameResult), C.c_int]),
    "pc_analyze_pending": (C.c_int, [C.c_void_p]),
    "pc_analyze_end": (C.c_int, [C.c_void_p]),
    "pc_analyze_set_halo": (C.c_int, [C.c_void_p, C.c_int]),
    "pc_analyze_track_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(BundleOpts), C.c_int, C.c_int]),
    "pc_analyze_track_seed": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(CameraState)]),
    "pc_mark": (C.c_int, [C.c_void_p, C.c_int]),
    "pc_elapsed_ms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "pc_synth_set_texture": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "pc_synth_render_rgb8": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.c_size_t]),
    "pc_device_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "pc_device_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pc_host_alloc_pinned": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "pc_host_free_pinned": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pc_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "pc_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "pc_timing_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "pc_timing_read": (C.c_int, [C.c_void_p, C.POINTER(KernelTimes), C.c_int]),
    "pc_mesh_set": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "pc_ray_cast": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(CameraState), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pc_solve_pnp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(BundleOpts), C.c_float, C.c_int, C.c_int, C.POINTER(CameraState), C.POINTER(BundleStats), C.POINTER(C.c_float)]),
    "pc_track_frame": (C.c_int, [C.c_void_p, C.POINTER(MatchSource), C.c_int, C.c_void_p, C.POINTER(CameraState), C.POINTER(BundleOpts), C.c_int, C.c_int, C.POINTER(CameraState), C.POINTER(BundleStats), C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "pc_ba_load": (C.c_int, [C.c_void_p, C.POINTER(BAProblem)]),
    "pc_ba_cost": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(BundleOpts), C.POINTER(C.c_float)]),
    "pc_ba_normal_equations": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(BundleOpts), C.c_void_p, C.c_void_p]),
    "pc_ba_read_cache": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "pc_ba_solve": (C.c_int, [C.c_void_p, C.POINTER(BundleOpts), C.c_void_p, C.POINTER(BundleStats), BA_ITER_CB, C.c_void_p]),
    "pc_ba_solve_step": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.POINTER(C.c_float)]),
    "pc_comm_unique_id": (C.c_int, [C.c_void_p]),
    "pc_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "pc_comm_destroy": (C.c_int, [C.c_void_p]),
    "pc_traj_allgather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "pc_ba_set_edge_shard": (C.c_int, [C.c_void_p, C.c_int]),
}


def load():
    """Loads the CUDA library.  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m polychase_b200.build` "
                          "(the product path is CUDA only; there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def default_gftt(**kw) -> GfttOpts:
    o = GfttOpts()
    load().pc_default_gftt_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def default_flow(**kw) -> FlowOpts:
    o = FlowOpts()
    load().pc_default_flow_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def default_bundle(**kw) -> BundleOpts:
    o = BundleOpts()
    load().pc_default_bundle_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def camera_state(K: dict, R: np.ndarray, t: np.ndarray) -> CameraState:
    """CameraState from intrinsics dict + rotation matrix/translation (object->camera)."""
    from .geometry import quat_from_matrix
    q = quat_from_matrix(np.asarray(R, np.float64))
    cs = CameraState(K["fx"], K["fy"], K["cx"], K["cy"], K["aspect_ratio"], K["width"], K["height"],
                     float(K["convention"]))
    cs.q[:] = [float(v) for v in q]
    cs.t[:] = [float(v) for v in t]
    cs.filled = 1.0
    return cs


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """Owns a pc_ctx.  Thin, typed wrappers; numpy in, numpy out."""

    def __init__(self, device: int = 0, max_width: int = 3840, max_height: int = 2160,
                 max_features: int = 16384, ring_frames: int = 0, pipeline_depth: int = 0):
        self.lib = load()
        lim = Limits(device, max_width, max_height, max_features, ring_frames, pipeline_depth)
        h = C.c_void_p()
        rc = self.lib.pc_create(C.byref(lim), C.byref(h))
        if rc != 0:
            raise PcError(rc, self.lib.pc_last_error(None).decode())
        self.h = h
        self.max_features = max_features

    def close(self):
        if getattr(self, "h", None):
            self.lib.pc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc: int):
        if rc != 0:
            raise PcError(rc, self.lib.pc_last_error(self.h).decode())

    # ---- frames ----
    def upload_rgb(self, frame_id: int, rgb: np.ndarray, flow: Optional[FlowOpts] = None):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        h, w, c = rgb.shape
        assert c == 3
        flow = flow or default_flow()
        self._chk(self.lib.pc_frame_upload_rgb8(self.h, frame_id, _ptr(rgb), w, h, w * 3, C.byref(flow)))

    def upload_gray(self, frame_id: int, gray: np.ndarray, flow: Optional[FlowOpts] = None):
        gray = np.ascontiguousarray(gray, np.uint8)
        h, w = gray.shape
        flow = flow or default_flow()
        self._chk(self.lib.pc_frame_upload_gray8(self.h, frame_id, _ptr(gray), w, h, w, C.byref(flow)))

    def frame_from_device(self, frame_id: int, dev_ptr: int, w: int, h: int, stride: int,
                          flow: Optional[FlowOpts] = None):
        flow = flow or default_flow()
        self._chk(self.lib.pc_frame_from_device_rgb8(self.h, frame_id, C.c_void_p(dev_ptr), w, h, stride,
                                                     C.byref(flow)))

    def release(self, frame_id: int):
        self._chk(self.lib.pc_frame_release(self.h, frame_id))

    def num_levels(self, frame_id: int) -> int:
        n = C.c_int()
        self._chk(self.lib.pc_frame_num_levels(self.h, frame_id, C.byref(n)))
        return n.value

    def read_level(self, frame_id: int, level: int) -> np.ndarray:
        w, h = C.c_int(), C.c_int()
        self._chk(self.lib.pc_frame_read_level(self.h, frame_id, level, None, 0, C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value), np.uint8)
        self._chk(self.lib.pc_frame_read_level(self.h, frame_id, level, _ptr(out), out.size, C.byref(w), C.byref(h)))
        return out

    # ---- detector ----
    def detect(self, frame_id: int, opts: Optional[GfttOpts] = None) -> np.ndarray:
        opts = opts or default_gftt()
        out = np.empty((self.max_features, 2), np.float32)
        n = C.c_int()
        self._chk(self.lib.pc_detect(self.h, frame_id, C.byref(opts), _ptr(out), self.max_features, C.byref(n)))
        return out[: n.value].copy()

    def min_eig_map(self, frame_id: int, w: int, h: int) -> np.ndarray:
        out = np.empty((h, w), np.float32)
        self._chk(self.lib.pc_min_eig_map(self.h, frame_id, _ptr(out), out.size))
        return out

    def set_keypoints(self, frame_id: int, kps: np.ndarray):
        kps = np.ascontiguousarray(kps, np.float32).reshape(-1, 2)
        self._chk(self.lib.pc_set_keypoints(self.h, frame_id, _ptr(kps), len(kps)))

    # ---- LK ----
    def lk_pair(self, frm: int, to: int, flow: Optional[FlowOpts] = None):
        flow = flow or default_flow()
        cap = self.max_features
        idx = np.empty((cap,), np.uint32)
        tgt = np.empty((cap, 2), np.float32)
        err = np.empty((cap,), np.float32)
        n = C.c_int()
        self._chk(self.lib.pc_lk_pair(self.h, frm, to, C.byref(flow), _ptr(idx), _ptr(tgt), _ptr(err), cap, C.byref(n)))
        k = n.value
        return idx[:k].copy(), tgt[:k].copy(), err[:k].copy()

    def lk_raw(self, frm: int, to: int, flow: Optional[FlowOpts] = None):
        flow = flow or default_flow()
        cap = self.max_features
        nxt = np.empty((cap, 2), np.float32)
        st = np.empty((cap,), np.uint8)
        err = np.empty((cap,), np.float32)
        n = C.c_int()
        self._chk(self.lib.pc_lk_raw(self.h, frm, to, C.byref(flow), _ptr(nxt), _ptr(st), _ptr(err), cap, C.byref(n)))
        k = n.value
        return nxt[:k].copy(), st[:k].copy(), err[:k].copy()

    # ---- streaming analyzer ----
    def analyze_begin(self, width: int, height: int, first_frame: int, num_frames: int,
                      gftt: Optional[GfttOpts] = None, flow: Optional[FlowOpts] = None):
        vi = VideoInfo(width, height, first_frame, num_frames)
        self._gftt = gftt or default_gftt()
        self._flow = flow or default_flow()
        self._chk(self.lib.pc_analyze_begin(self.h, C.byref(vi), C.byref(self._gftt), C.byref(self._flow)))

    def analyze_push(self, frame_id: int, rgb, stride: int = 0, mem_kind: int = PC_MEM_HOST):
        if isinstance(rgb, np.ndarray):
            assert rgb.flags["C_CONTIGUOUS"] and rgb.dtype == np.uint8
            stride = stride or rgb.shape[1] * 3
            p = _ptr(rgb)
        else:
            p = C.c_void_p(int(rgb))
        self._chk(self.lib.pc_analyze_push_frame(self.h, frame_id, p, stride, mem_kind))

    def analyze_preset_keypoints(self, frame_id: int, kps: np.ndarray):
        kps = np.ascontiguousarray(kps, np.float32).reshape(-1, 2)
        self._chk(self.lib.pc_analyze_preset_keypoints(self.h, frame_id, _ptr(kps), len(kps)))

    def analyze_pending(self) -> int:
        return self.lib.pc_analyze_pending(self.h)

    def analyze_pop(self, download: bool = True, copy: bool = True):
        """Returns dict(frame_id, keypoints, pairs=[(from, to, idx, tgt, err)]) or counts only."""
        r = FrameResult()
        self._chk(self.lib.pc_analyze_pop(self.h, C.byref(r), 1 if download else 0))
        out = {"frame_id": r.frame_id, "num_keypoints": r.num_keypoints, "pairs": []}
        if download:
            k = np.ctypeslib.as_array(r.keypoints, shape=(r.num_keypoints, 2)) if r.num_keypoints else np.zeros((0, 2), np.float32)
            out["keypoints"] = k.copy() if copy else k
        for i in range(r.num_pairs):
            p = r.pairs[i]
            if download and p.rows:
                idx = np.ctypeslib.as_array(p.src_kps_indices, shape=(p.rows,))
                tgt = np.ctypeslib.as_array(p.tgt_kps, shape=(p.rows, 2))
                err = np.ctypeslib.as_array(p.flow_errors, shape=(p.rows,))
                if copy:
                    idx, tgt, err = idx.copy(), tgt.copy(), err.copy()
            elif download:
                idx, tgt, err = np.zeros((0,), np.uint32), np.zeros((0, 2), np.float32), np.zeros((0,), np.float32)
            else:
                idx = tgt = err = None
            out["pairs"].append((p.image_id_from, p.image_id_to, p.rows, idx, tgt, err))
        out["tracked"] = r.tracked
        if r.tracked:
            out["camera"] = CameraState.from_buffer_copy(r.camera)
            out["num_matches"] = r.num_matches
            out["inlier_ratio"] = float(r.inlier_ratio)
            out["stats"] = BundleStats.from_buffer_copy(r.stats)
        return out

    def analyze_track_begin(self, model: np.ndarray, opts: Optional[BundleOpts] = None, opt_f: bool = False,
                            opt_pp: bool = False):
        """Chains a forward TrackSequence behind the analyzer (call after analyze_begin + mesh_set)."""
        model = np.ascontiguousarray(model, np.float32).reshape(16)
        opts = opts or default_bundle()
        self._chk(self.lib.pc_analyze_track_begin(self.h, _ptr(model), C.byref(opts), int(opt_f), int(opt_pp)))

    def analyze_track_seed(self, frame_id: int, cam: CameraState):
        self._chk(self.lib.pc_analyze_track_seed(self.h, frame_id, C.byref(cam)))

    def analyze_set_halo(self, n: int):
        self._chk(self.lib.pc_analyze_set_halo(self.h, n))

    def mark(self, slot: int):
        self._chk(self.lib.pc_mark(self.h, slot))

    def elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        self._chk(self.lib.pc_elapsed_ms(self.h, a, b, C.byref(ms)))
        return float(ms.value)

    def analyze_end(self):
        self._chk(self.lib.pc_analyze_end(self.h))

    # ---- misc ----
    def synchronize(self):
        self._chk(self.lib.pc_synchronize(self.h))

    def kernel_launches(self) -> int:
        return int(self.lib.pc_kernel_launches(self.h))

    def synth_set_texture(self, tex: np.ndarray):
        tex = np.ascontiguousarray(tex, np.uint8)
        self._chk(self.lib.pc_synth_set_texture(self.h, _ptr(tex), tex.shape[1], tex.shape[0]))

    def synth_render(self, H: np.ndarray, dev_ptr: int, stride: int):
        Hc = (C.c_double * 9)(*[float(v) for v in np.asarray(H, np.float64).ravel()])
        self._chk(self.lib.pc_synth_render_rgb8(self.h, Hc, C.c_void_p(dev_ptr), stride))

    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._chk(self.lib.pc_device_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def device_free(self, p: int):
        self._chk(self.lib.pc_device_free(self.h, C.c_void_p(p)))

    def pinned_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._chk(self.lib.pc_host_alloc_pinned(self.h, nbytes, C.byref(p)))
        return p.value

    def pinned_free(self, p: int):
        self._chk(self.lib.pc_host_free_pinned(self.h, C.c_void_p(p)))

    def d2h(self, dev_ptr: int, nbytes: int) -> np.ndarray:
        out = np.empty((nbytes,), np.uint8)
        self._chk(self.lib.pc_memcpy_d2h(self.h, _ptr(out), C.c_void_p(dev_ptr), nbytes))
        return out

    def h2d(self, dev_ptr: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        self._chk(self.lib.pc_memcpy_h2d(self.h, C.c_void_p(dev_ptr), _ptr(arr), arr.nbytes))

    def timing_enable(self, on: bool = True):
        self._chk(self.lib.pc_timing_enable(self.h, 1 if on else 0))

    def timing_read(self, reset: bool = True) -> dict:
        t = KernelTimes()
        self._chk(self.lib.pc_timing_read(self.h, C.byref(t), 1 if reset else 0))
        return {n: getattr(t, n) for n, _ in KernelTimes._fields_}


# ---- track / refine wrappers (methods added to Context) --------------------------------------
def _ctx_mesh_set(self, verts: np.ndarray, tris: np.ndarray, mask_bits: Optional[np.ndarray] = None):
    verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
    tris = np.ascontiguousarray(tris, np.uint32).reshape(-1, 3)
    mb = None if mask_bits is None or len(mask_bits) == 0 else np.ascontiguousarray(mask_bits, np.uint32)
    self._chk(self.lib.pc_mesh_set(self.h, _ptr(verts), len(verts), _ptr(tris), len(tris), _ptr(mb),
                                   0 if mb is None else len(mb)))


def _ctx_ray_cast(self, model: np.ndarray, cam: CameraState, pos: np.ndarray, check_mask: bool = True):
    model = np.ascontiguousarray(model, np.float32).reshape(16)
    pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 2)
    n = len(pos)
    hit = np.zeros(n, np.uint8)
    P = np.zeros((n, 3), np.float32)
    prim = np.zeros(n, np.uint32)
    uv = np.zeros((n, 2), np.float32)
    t = np.zeros(n, np.float32)
    self._chk(self.lib.pc_ray_cast(self.h, _ptr(model), C.byref(cam), _ptr(pos), n, 1 if check_mask else 0,
                                   _ptr(hit), _ptr(P), _ptr(prim), _ptr(uv), _ptr(t)))
    return hit.astype(bool), P, prim, uv, t


def _ctx_solve_pnp(self, X, x, cam: CameraState, opts: Optional[BundleOpts] = None, weights=None,
                   max_inlier_error: float = 12.0, opt_f: bool = False, opt_pp: bool = False):
    X = np.ascontiguousarray(X, np.float32).reshape(-1, 3)
    x = np.ascontiguousarray(x, np.float32).reshape(-1, 2)
    w = None if weights is None else np.ascontiguousarray(weights, np.float32)
    opts = opts or default_bundle()
    out = CameraState.from_buffer_copy(cam)
    st = BundleStats()
    inl = C.c_float()
    self._chk(self.lib.pc_solve_pnp(self.h, _ptr(X), _ptr(x), _ptr(w), len(X), C.byref(opts), max_inlier_error,
                                    int(opt_f), int(opt_pp), C.byref(out), C.byref(st), C.byref(inl)))
    return out, st, float(inl.value)


def _ctx_track_frame(self, sources, model, init: CameraState, opts: Optional[BundleOpts] = None,
                     opt_f: bool = False, opt_pp: bool = False):
    """sources: list of (CameraState, keypoints (nk,2), src_idx (rows,), tgt (rows,2))."""
    opts = opts or default_bundle()
    model = np.ascontiguousarray(model, np.float32).reshape(16)
    arr = (MatchSource * max(len(sources), 1))()
    keep = []
    for i, (cam, kps, idx, tgt) in enumerate(sources):
        kps = np.ascontiguousarray(kps, np.float32).reshape(-1, 2)
        idx = np.ascontiguousarray(idx, np.uint32)
        tgt = np.ascontiguousarray(tgt, np.float32).reshape(-1, 2)
        keep += [kps, idx, tgt]
        arr[i] = MatchSource(cam, kps.ctypes.data, len(kps), idx.ctypes.data, tgt.ctypes.data, len(idx))
    out = CameraState()
    st = BundleStats()
    inl = C.c_float()
    nm = C.c_int()
    self._chk(self.lib.pc_track_frame(self.h, arr, len(sources), _ptr(model), C.byref(init), C.byref(opts),
                                      int(opt_f), int(opt_pp), C.byref(out), C.byref(st), C.byref(inl), C.byref(nm)))
    return out, st, float(inl.value), nm.value


def _ctx_ba_load(self, keypoints, edges, model, opt_f: bool = False, opt_pp: bool = False):
    """keypoints: list of (n_i,2) per frame; edges: list of (src, tgt, src_idx, tgt_kps)."""
    offs = np.concatenate([[0], np.cumsum([len(k) for k in keypoints])]).astype(np.int32)
    kps = (np.concatenate([np.asarray(k, np.float32).reshape(-1, 2) for k in keypoints])
           if offs[-1] else np.zeros((0, 2), np.float32))
    kps = np.ascontiguousarray(kps, np.float32)
    ed = (BAEdge * max(len(edges), 1))()
    idx_all, tgt_all = [], []
    row = 0
    for i, (s, t, idx, tgt) in enumerate(edges):
        ed[i] = BAEdge(int(s), int(t), row, len(idx))
        idx_all.append(np.asarray(idx, np.uint32))
        tgt_all.append(np.asarray(tgt, np.float32).reshape(-1, 2))
        row += len(idx)
    idx_all = np.ascontiguousarray(np.concatenate(idx_all) if idx_all else np.zeros(0, np.uint32), np.uint32)
    tgt_all = np.ascontiguousarray(np.concatenate(tgt_all) if tgt_all else np.zeros((0, 2), np.float32), np.float32)
    pr = BAProblem()
    pr.num_frames = len(keypoints)
    pr.kp_offsets = offs.ctypes.data
    pr.keypoints = kps.ctypes.data
    pr.num_edges = len(edges)
    pr.edges = C.addressof(ed)
    pr.src_kps_indices = idx_all.ctypes.data
    pr.tgt_kps = tgt_all.ctypes.data
    pr.model[:] = [float(v) for v in np.asarray(model, np.float32).reshape(16)]
    pr.optimize_focal_length = int(opt_f)
    pr.optimize_principal_point = int(opt_pp)
    self._chk(self.lib.pc_ba_load(self.h, C.byref(pr)))
    self._ba_nf = len(keypoints)
    self._ba_p = 9 if (opt_f or opt_pp) else 6


def _traj_array(traj):
    arr = (CameraState * len(traj))()
    for i, c in enumerate(traj):
        arr[i] = c
    return arr


def _ctx_ba_cost(self, traj, opts: Optional[BundleOpts] = None) -> float:
    opts = opts or default_bundle()
    arr = _traj_array(traj)
    cost = C.c_float()
    self._chk(self.lib.pc_ba_cost(self.h, arr, C.byref(opts), C.byref(cost)))
    return float(cost.value)


def _ctx_ba_normal_equations(self, traj, opts: Optional[BundleOpts] = None):
    opts = opts or default_bundle()
    arr = _traj_array(traj)
    nf, p = self._ba_nf, self._ba_p
    band = np.zeros((nf, 9, p, p), np.float32)
    jtr = np.zeros((nf * p,), np.float32)
    self._chk(self.lib.pc_ba_normal_equations(self.h, arr, C.byref(opts), _ptr(band), _ptr(jtr)))
    return band, jtr


def _ctx_ba_solve(self, traj, opts: Optional[BundleOpts] = None, callback=None):
    opts = opts or default_bundle()
    arr = _traj_array(traj)
    st = BundleStats()
    if callback is not None:
        cb = BA_ITER_CB(lambda s, u: 1 if callback(s.contents) else 0)
    else:
        cb = C.cast(None, BA_ITER_CB)
    self._chk(self.lib.pc_ba_solve(self.h, C.byref(opts), arr, C.byref(st), cb, None))
    return [CameraState.from_buffer_copy(arr[i]) for i in range(len(traj))], st


def _ctx_ba_read_cache(self, n: int) -> np.ndarray:
    out = np.zeros(n, np.uint32)
    self._chk(self.lib.pc_ba_read_cache(self.h, _ptr(out), n))
    return out


def band_to_dense(band: np.ndarray) -> np.ndarray:
    """Lower-triangular dense matrix from the (nf, 9, p, p) band layout of pc_ba_normal_equations."""
    nf, nb, p, _ = band.shape
    A = np.zeros((nf * p, nf * p), band.dtype)
    for i in range(nf):
        for k in range(nb):
            j = i - k
            if j < 0:
                continue
            blk = band[i, k]
            A[i * p:(i + 1) * p, j * p:(j + 1) * p] = np.tril(blk) if k == 0 else blk
    return A


def _ctx_ba_solve_step(self, lam: float):
    """step = -(A with diag * (1 + lam))^-1 Jtr of the last ba_normal_equations (K14 on its own)."""
    out = np.zeros(self._ba_nf * self._ba_p, np.float32)
    sn = C.c_float()
    self._chk(self.lib.pc_ba_solve_step(self.h, float(lam), _ptr(out), C.byref(sn)))
    return out, float(sn.value)


COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the C ABI (rank 0); carry the bytes to the other ranks."""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    rc = load().pc_comm_unique_id(buf)
    if rc != 0:
        raise PcError(rc, load().pc_last_error(None).decode())
    return bytes(buf)


def _ctx_comm_init(self, world: int, rank: int, uid: bytes):
    buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(uid)
    self._chk(self.lib.pc_comm_init(self.h, world, rank, buf))
    self._comm = (world, rank)


def _ctx_comm_destroy(self):
    self._chk(self.lib.pc_comm_destroy(self.h))


def _ctx_traj_allgather(self, local, counts):
    """local: list of CameraState (this rank's segment); counts: every rank's segment length."""
    arr = _traj_array(local) if len(local) else (CameraState * 1)()
    cnt = (C.c_int * len(counts))(*[int(x) for x in counts])
    out = (CameraState * max(int(sum(counts)), 1))()
    self._chk(self.lib.pc_traj_allgather(self.h, arr, len(local), cnt, out))
    return [CameraState.from_buffer_copy(out[i]) for i in range(int(sum(counts)))]


def _ctx_ba_set_edge_shard(self, on: bool = True):
    self._chk(self.lib.pc_ba_set_edge_shard(self.h, 1 if on else 0))


Context.mesh_set = _ctx_mesh_set
Context.ba_solve_step = _ctx_ba_solve_step
Context.comm_init = _ctx_comm_init
Context.comm_destroy = _ctx_comm_destroy
Context.traj_allgather = _ctx_traj_allgather
Context.ba_set_edge_shard = _ctx_ba_set_edge_shard
Context.ray_cast = _ctx_ray_cast
Context.solve_pnp = _ctx_solve_pnp
Context.track_frame = _ctx_track_frame
Context.ba_load = _ctx_ba_load
Context.ba_cost = _ctx_ba_cost
Context.ba_normal_equations = _ctx_ba_normal_equations
Context.ba_solve = _ctx_ba_solve
Context.ba_read_cache = _ctx_ba_read_cache
