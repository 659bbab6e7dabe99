"""The reference detector's own logic around the min-eigenvalue map -- TEST INFRASTRUCTURE.

Restates /root/reference/cpp/feature_detection/gftt.cc:38-192 (C implementation in
oracle/restate.c::orc_gftt_select; no cv2 involved, so it is deterministic anywhere)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import restate


def num_pyramid_levels(w: int, h: int, win: int = 10, max_level: int = 3) -> int:
    """cv::buildOpticalFlowPyramid: stop after level L when the next size <= winSize."""
    levels = 0
    for L in range(max_level + 1):
        levels = L + 1
        w, h = (w + 1) // 2, (h + 1) // 2
        if w <= win or h <= win:
            break
    return levels


def gftt_from_eig(eig: np.ndarray, quality_level: float = 0.01, min_distance: float = 5.0,
                  max_corners: int = 0, grid_rows: int = 4, grid_cols: int = 4) -> np.ndarray:
    lib = restate.lib()
    lib.orc_gftt_select.restype = C.c_int
    lib.orc_gftt_select.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_int]
    e = np.array(eig, np.float32, copy=True, order="C")
    h, w = e.shape
    cap = w * h // 2 + 16
    out = np.empty((cap, 2), np.float32)
    n = lib.orc_gftt_select(e.ctypes.data_as(C.c_void_p), w, h, quality_level, min_distance, max_corners,
                            grid_rows, grid_cols, out.ctypes.data_as(C.c_void_p), cap)
    assert n >= 0
    return out[:n].copy()


def detect(gray: np.ndarray, mode: int = 1, **kw) -> np.ndarray:
    """Full restated detector on a gray image (mode: see restate.min_eig)."""
    return gftt_from_eig(restate.min_eig(gray, mode), **kw)
