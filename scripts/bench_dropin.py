"""The reference's own entry point for Analyze Video, timed end to end: polychase_core.OpticalFlowThread
(frame request / provide_frame hand-off, polychase_pybind.cc:29-348, opticalflow_thread.h) writing the
reference's SQLite database.  Frames are host numpy arrays (what the Blender addon hands over).

    python scripts/bench_dropin.py [--config 4k] [--frames 64]

Prints one JSON line: frames/s and directed pairs/s of the whole call, database size.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {"4k": (3840, 2160, 8000), "1080p": (1920, 1080, 4000), "720p": (1280, 720, 2000)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="4k", choices=sorted(CONFIGS))
    ap.add_argument("--frames", type=int, default=64)
    args = ap.parse_args()
    from oracle import synth          # input generator only
    from polychase_b200 import capi, polychase_core as core

    w, h, max_corners = CONFIGS[args.config]
    F = args.frames
    # render the clip on the GPU, keep it as host arrays
    ctx = capi.Context(device=0, max_width=w, max_height=h, max_features=1024)
    ctx.synth_set_texture(synth.make_texture(w, h, seed=0))
    K = synth.intrinsics(w, h)
    scale = synth.plane_scale(w, 4.0)
    Rs, ts = synth.camera_path(F, 4.0, 0)
    dev = ctx.device_alloc(w * h * 3)
    frames = {}
    for i in range(F):
        ctx.synth_render(synth.homography(K, Rs[i], ts[i], w, h, scale), dev, w * 3)
        ctx.synchronize()
        a = np.empty((h, w, 3), np.uint8)
        ctx.lib.pc_memcpy_d2h(ctx.h, a.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(dev), a.nbytes)
        frames[i + 1] = a
    ctx.device_free(dev)
    ctx.close()

    tmp = tempfile.mkdtemp()
    dbp = os.path.join(tmp, "clip.db")
    go = core.GFTTOptions()
    go.max_corners = max_corners
    errors = []
    t0 = time.perf_counter()
    th = core.OpticalFlowThread(core.VideoInfo(w, h, 1, F), dbp, go)
    done = False
    provide_s = 0.0
    first_provide = None
    while not done:
        m = th.try_pop()
        if m is None:
            time.sleep(0.0002)
            continue
        if isinstance(m, bool):
            done = True
        elif isinstance(m, core.OpticalFlowRequest):
            t1 = time.perf_counter()
            th.provide_frame(m.frame_id, frames[m.frame_id])
            t2 = time.perf_counter()
            if first_provide is None:
                first_provide = (t1 - t0, t2 - t1)
            else:
                provide_s += t2 - t1
        elif isinstance(m, core.CppException):
            errors.append(m.what())
    th.join()
    dt = time.perf_counter() - t0
    pairs = 8 * F - 30
    print(json.dumps({"metric": "OpticalFlowThread (Analyze Video) wall clock at %s" % args.config, "frames": F,
                      "directed_pairs": pairs, "wall_s": dt, "frames_per_s": F / dt, "pairs_per_s": pairs / dt,
                      "db_bytes": os.path.getsize(dbp), "errors": errors,
                      "first_request_after_s": first_provide[0], "first_provide_s": first_provide[1],
                      "other_provides_total_s": provide_s}))


if __name__ == "__main__":
    main()
