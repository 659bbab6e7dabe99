"""CPU restatement of the reference's Analyze pass around OpenCV (cv2) -- TEST
INFRASTRUCTURE and bench.py's CPU baseline.

Follows GenerateOpticalFlowDatabase, /root/reference/cpp/opticalflow.cc:209-321, including
its redundancy: every frame is re-converted to gray and re-pyramided once per pair it is
the target of (opticalflow.cc:298-302), and up to 4 pairs are in flight
(max_allowed_parallelism 4, opticalflow.cc:270-271).  The cv2 Python binding cannot take a
prebuilt pyramid, so cv::calcOpticalFlowPyrLK rebuilds the *source* frame's pyramid inside
each call too (9 extra pyrDown chains per frame relative to the C++ reference; noted in
DESIGN.md section "CPU baseline").
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import cvref

IMAGE_SKIPS = (-8, -4, -2, -1, 1, 2, 4, 8)   # opticalflow.cc:76-77


def analyze_frame(frame_accessor, frame_id1: int, first: int, last_excl: int, pool, gftt_kw: dict,
                  flow_kw: dict, known_kps=None):
    """One iteration of the outer loop (opticalflow.cc:237-316).  Returns (keypoints,
    [(from, to, src_idx, tgt, err), ...])."""
    frame1 = frame_accessor(frame_id1)
    gray1 = cvref.rgb2gray(frame1)                                   # :259
    if known_kps is not None and len(known_kps):                     # ReadOrGenerateKeypoints :168-178
        kps = known_kps
    else:
        kps, _ = cvref.gftt(gray1, **gftt_kw)                        # :154-166

    def one_pair(skip):
        frame_id2 = frame_id1 + skip
        if frame_id2 < first or frame_id2 >= last_excl:             # :281-284
            return None
        gray2 = cvref.rgb2gray(frame_accessor(frame_id2))            # :298
        nxt, st, err = cvref.lk(gray1, gray2, kps, **flow_kw)        # :119-125 (pyramids inside)
        ok = st == 1                                                 # :139-147
        return (frame_id1, frame_id2, np.nonzero(ok)[0].astype(np.uint32), nxt[ok], err[ok])

    rows = [r for r in pool.map(one_pair, IMAGE_SKIPS) if r is not None]
    return kps, rows


def analyze_clip(frame_accessor, first: int, num: int, gftt_kw=None, flow_kw=None, threads: int = 4):
    gftt_kw = gftt_kw or {}
    flow_kw = flow_kw or {}
    out_kps, out_rows = {}, []
    with ThreadPoolExecutor(max_workers=threads) as pool:
        for f in range(first, first + num):
            kps, rows = analyze_frame(frame_accessor, f, first, first + num, pool, gftt_kw, flow_kw)
            out_kps[f] = kps
            out_rows += rows
    return out_kps, out_rows
