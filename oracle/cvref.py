"""OpenCV itself (cv2) for the stages the reference delegates to OpenCV -- TEST
INFRASTRUCTURE.

The reference calls (all third-party, source absent from /root/reference):
  cv::cvtColor(RGB2GRAY)              cpp/opticalflow.cc:259,298
  cv::buildOpticalFlowPyramid         cpp/opticalflow.cc:184-186
  cv::calcOpticalFlowPyrLK            cpp/opticalflow.cc:119-125
  cv::cornerMinEigenVal/minMaxLoc/threshold/dilate   cpp/feature_detection/gftt.cc:35-36,62-65,70
Reference pin: vcpkg tag 2025.06.13 (opencv4 4.11.x, no IPP).  Here: the cv2 wheel.

`pin()` fixes the cv2 runtime configuration the goldens are generated under.
"""
from __future__ import annotations

import cv2
import numpy as np


def pin(threads: int = 1) -> None:
    cv2.setUseOptimized(True)
    try:
        cv2.ipp.setUseIPP(False)
    except Exception:  # pragma: no cover
        pass
    cv2.setNumThreads(threads)


def build_info_digest() -> str:
    info = cv2.getBuildInformation()
    keep = [l.strip() for l in info.splitlines()
            if any(k in l for k in ("Version control", "CPU/HW features", "Baseline",
                                    "Dispatched code", "requested", "Parallel framework",
                                    "Intel IPP"))]
    return f"cv2 {cv2.__version__}; " + "; ".join(keep)


def rgb2gray(rgb: np.ndarray) -> np.ndarray:
    return cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)


def pyramid(gray: np.ndarray, win: int = 10, max_level: int = 3):
    """Returns (levels, derivs): unpadded u8 level images and int16 (h,w,2) Scharr
    derivative images, as cv::buildOpticalFlowPyramid lays them out."""
    n, pyr = cv2.buildOpticalFlowPyramid(gray, (win, win), max_level, withDerivatives=True)
    levels = [np.ascontiguousarray(pyr[2 * i]) for i in range(n + 1)]
    derivs = [np.ascontiguousarray(pyr[2 * i + 1]) for i in range(n + 1)]
    return levels, derivs


def lk(gray1: np.ndarray, gray2: np.ndarray, pts: np.ndarray, win: int = 10,
       max_level: int = 3, iters: int = 30, eps: float = 0.01, min_eig: float = 1e-4):
    """cv::calcOpticalFlowPyrLK with the reference's arguments (opticalflow.cc:119-125).
    Returns next (N,2) f32, status (N,) u8, err (N,) f32."""
    if len(pts) == 0:
        return (np.zeros((0, 2), np.float32), np.zeros((0,), np.uint8),
                np.zeros((0,), np.float32))
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 1, 2)
    nxt, st, err = cv2.calcOpticalFlowPyrLK(
        gray1, gray2, p, None, winSize=(win, win), maxLevel=max_level,
        criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, iters, eps),
        flags=0, minEigThreshold=min_eig)
    return nxt.reshape(-1, 2), st.reshape(-1), err.reshape(-1)


def min_eig(gray: np.ndarray, block: int = 3, ksize: int = 3) -> np.ndarray:
    return cv2.cornerMinEigenVal(gray, block, ksize=ksize)


def gftt(gray: np.ndarray, quality_level: float = 0.01, min_distance: float = 5.0,
         block_size: int = 3, gradient_size: int = 3, max_corners: int = 0,
         grid_rows: int = 4, grid_cols: int = 4):
    """The reference's GoodFeaturesToTrack (cpp/feature_detection/gftt.cc:14-192)
    restated around the same cv2 calls it makes.  Returns (N,2) f32 of integer (x,y)
    in the reference's output order, plus the eig map (before thresholding)."""
    eig = cv2.cornerMinEigenVal(gray, block_size, ksize=gradient_size)
    eig_raw = eig.copy()
    h, w = gray.shape
    grid_rows = max(1, grid_rows)
    grid_cols = max(1, grid_cols)
    bh = (h + grid_rows - 1) // grid_rows          # gftt.cc:42
    bw = (w + grid_cols - 1) // grid_cols          # gftt.cc:43
    for gy in range(grid_rows):                    # gftt.cc:47-66
        for gx in range(grid_cols):
            y0, x0 = gy * bh, gx * bw
            y1, x1 = min(y0 + bh, h), min(x0 + bw, w)
            if y1 <= y0 or x1 <= x0:
                continue
            blk = eig[y0:y1, x0:x1]
            _, max_val, _, _ = cv2.minMaxLoc(blk)
            cv2.threshold(blk, max_val * quality_level, 0, cv2.THRESH_TOZERO, dst=blk)
    tmp = cv2.dilate(eig, None)                    # gftt.cc:70
    inner = np.zeros_like(eig, bool)
    inner[1:h - 1, 1:w - 1] = True                 # gftt.cc:76-80
    cand = inner & (eig != 0) & (eig == tmp)       # gftt.cc:83
    ys, xs = np.nonzero(cand)
    vals = eig[ys, xs]
    addr = ys.astype(np.int64) * w + xs
    # sort by value desc, then address desc (gftt.cc:7-12,98)
    order = np.lexsort((-addr, -vals.astype(np.float64)))
    xs, ys = xs[order], ys[order]
    out = greedy_min_distance(xs, ys, w, h, min_distance, max_corners)
    return out, eig_raw


def greedy_min_distance(xs, ys, w, h, min_distance, max_corners):
    """gftt.cc:100-190 -- sequential greedy suppression on the sorted candidates."""
    total = len(xs)
    if min_distance >= 1:
        cell = int(np.rint(min_distance))          # cvRound, gftt.cc:105
        gw = (w + cell - 1) // cell
        gh = (h + cell - 1) // cell
        grid = {}
        md2 = min_distance * min_distance
        out = []
        for i in range(total):
            x, y = int(xs[i]), int(ys[i])
            xc, yc = x // cell, y // cell
            x1, y1 = max(xc - 1, 0), max(yc - 1, 0)
            x2, y2 = min(xc + 1, gw - 1), min(yc + 1, gh - 1)
            good = True
            for yy in range(y1, y2 + 1):
                for xx in range(x1, x2 + 1):
                    for (px, py) in grid.get((yy, xx), ()):
                        dx, dy = x - px, y - py
                        if dx * dx + dy * dy < md2:
                            good = False
                            break
                    if not good:
                        break
                if not good:
                    break
            if good:
                grid.setdefault((yc, xc), []).append((x, y))
                out.append((x, y))
                if max_corners > 0 and len(out) == max_corners:
                    break
        return np.asarray(out, np.float32).reshape(-1, 2)
    n = total if max_corners <= 0 else min(total, max_corners)
    return np.stack([xs[:n], ys[:n]], axis=1).astype(np.float32).reshape(-1, 2)
