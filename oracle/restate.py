"""ctypes front-end of oracle/restate.c -- TEST INFRASTRUCTURE (see that file's header)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build as _build

_lib = None


class _Level(C.Structure):
    _fields_ = [("img", C.c_void_p), ("deriv", C.c_void_p), ("w", C.c_int), ("h", C.c_int)]


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.orc_rgb2gray.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p]
        _lib.orc_pyrdown.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.orc_scharr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.orc_min_eig.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.orc_lk.argtypes = [C.POINTER(_Level), C.POINTER(_Level), C.c_int, C.c_void_p, C.c_int,
                                C.c_int, C.c_int, C.c_double, C.c_double,
                                C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def rgb2gray(rgb: np.ndarray) -> np.ndarray:
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, _ = rgb.shape
    out = np.empty((h, w), np.uint8)
    lib().orc_rgb2gray(_p(rgb), w, h, w * 3, _p(out))
    return out


def pyrdown(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().orc_pyrdown(_p(img), w, h, _p(out))
    return out


def pyramid(gray: np.ndarray, max_level: int = 3):
    levels = [np.ascontiguousarray(gray, np.uint8)]
    for _ in range(max_level):
        levels.append(pyrdown(levels[-1]))
    return levels


def scharr(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty((h, w, 2), np.int16)
    lib().orc_scharr(_p(img), w, h, _p(out))
    return out


def min_eig(gray: np.ndarray, mode: int = 1) -> np.ndarray:
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    out = np.empty((h, w), np.float32)
    lib().orc_min_eig(_p(gray), w, h, mode, _p(out))
    return out


def lk(levels1, levels2, pts: np.ndarray, win: int = 10, iters: int = 30, eps: float = 0.01,
       min_eig_thr: float = 1e-4):
    """Pyramidal LK on prebuilt level lists (index 0 = full res)."""
    n = len(pts)
    nl = len(levels1)
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    derivs = [scharr(l) for l in levels1]
    A = (_Level * nl)()
    B = (_Level * nl)()
    keep = []
    for i in range(nl):
        a = np.ascontiguousarray(levels1[i], np.uint8)
        b = np.ascontiguousarray(levels2[i], np.uint8)
        keep += [a, b]
        A[i] = _Level(a.ctypes.data, derivs[i].ctypes.data, a.shape[1], a.shape[0])
        B[i] = _Level(b.ctypes.data, None, b.shape[1], b.shape[0])
    nxt = np.zeros((n, 2), np.float32)
    st = np.zeros((n,), np.uint8)
    err = np.zeros((n,), np.float32)
    if n:
        lib().orc_lk(A, B, nl, _p(pts), n, win, iters, eps, min_eig_thr, _p(nxt), _p(st), _p(err))
    return nxt, st, err
