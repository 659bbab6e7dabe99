"""Does the CUDA detector's one known arithmetic deviation ever change a feature list at 4K?

The kernel forms every 3x3 box sum independently (oracle/restate.c::orc_min_eig mode 3) while
cv::boxFilter carries a running double column sum from row 0 (mode 1, bit-equal to cv2).  This
script runs the detector logic of gftt.cc:38-192 (oracle/gftt.py) on both eig maps for N frames
of the bench clip (4K, max_corners 8000, seed 0) and reports: differing eig pixels, how far the
differing pixels sit below their grid cell's threshold, and whether any feature list differs.

    python scripts/detector_4k_divergence.py [--frames 64] [--config 4k] > profiles/r2_detector_4k_divergence.json
CPU only (oracle = test infrastructure); nothing here is on the product path."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {"4k": (3840, 2160, 8000), "1080p": (1920, 1080, 4000), "720p": (1280, 720, 2000)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--config", default="4k")
    ap.add_argument("--stride", type=int, default=1)
    args = ap.parse_args()
    from oracle import gftt as ogftt
    from oracle import restate, synth
    w, h, mc = CONFIGS[args.config]
    clip = synth.Clip(w, h, args.frames * args.stride, seed=0)
    out = dict(config=args.config, frames=args.frames, max_corners=mc, differing_pixels=[], lists_differ=0,
               min_margin_below_threshold=None, unlimited_lists_differ=0)
    worst = None
    for i in range(args.frames):
        k = i * args.stride
        g = clip.gray(k)
        e1 = restate.min_eig(g, 1)
        e3 = restate.min_eig(g, 3)
        diff = np.nonzero(e1 != e3)
        nd = len(diff[0])
        out["differing_pixels"].append(nd)
        # per-cell threshold (gftt.cc:61-65): 0.01 * max of the 4x4 cell
        bh, bw = (h + 3) // 4, (w + 3) // 4
        for y, x in zip(*diff):
            cy, cx = y // bh, x // bw
            cell = e1[cy * bh:(cy + 1) * bh, cx * bw:(cx + 1) * bw]
            thr = 0.01 * float(cell.max())
            ratio = max(float(e1[y, x]), float(e3[y, x])) / thr if thr > 0 else float("inf")
            worst = ratio if worst is None else max(worst, ratio)
        k1 = ogftt.gftt_from_eig(e1.copy(), max_corners=mc)
        k3 = ogftt.gftt_from_eig(e3.copy(), max_corners=mc)
        if not np.array_equal(k1, k3):
            out["lists_differ"] += 1
        u1 = ogftt.gftt_from_eig(e1.copy(), max_corners=0)
        u3 = ogftt.gftt_from_eig(e3.copy(), max_corners=0)
        if not np.array_equal(u1, u3):
            out["unlimited_lists_differ"] += 1
        print(f"frame {k}: {nd} differing eig px, list equal {np.array_equal(k1, k3)}, unlimited equal "
              f"{np.array_equal(u1, u3)} ({len(u1)} corners)", file=sys.stderr, flush=True)
    out["total_differing_pixels"] = int(sum(out["differing_pixels"]))
    out["px_per_mpx"] = out["total_differing_pixels"] / (args.frames * w * h / 1e6)
    out["max_value_over_cell_threshold_at_differing_pixels"] = worst
    print(json.dumps(out))


if __name__ == "__main__":
    main()
