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
    ap.add_argument("--no-refine", action="store_true")
    args = ap.parse_args()
    from polychase_b200 import synth  # input generator
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
    out_lines = []
    out_lines.append(json.dumps({"metric": "OpticalFlowThread (Analyze Video) wall clock at %s" % args.config, "frames": F,
                      "directed_pairs": pairs, "wall_s": dt, "frames_per_s": F / dt, "pairs_per_s": pairs / dt,
                      "db_bytes": os.path.getsize(dbp), "errors": errors,
                      "first_request_after_s": first_provide[0], "first_provide_s": first_provide[1],
                      "other_provides_total_s": provide_s}))

    # ---- TrackerThread (Track Sequence, tracker.cc:194-213) over the same database ---------------
    def pump(thread, on_msg):
        while True:
            m = thread.try_pop()
            if m is None:
                time.sleep(0.0002)
                continue
            if isinstance(m, bool):
                return
            on_msg(m)

    verts, tris = synth.plane_mesh(w, h, scale)
    mesh = core.AcceleratedMesh(verts, tris)
    intr = core.CameraIntrinsics(K["fx"], K["fy"], K["cx"], K["cy"], 1.0, w, h, core.CameraConvention.OpenCV)
    view = np.eye(4, dtype=np.float32)
    view[:3, :3] = Rs[0]
    view[:3, 3] = ts[0]
    scene = core.SceneTransformations(np.eye(4, dtype=np.float32), view, intr)
    bo = core.BundleOptions()
    bo.loss_type = core.LossType.Cauchy
    results = []
    t0 = time.perf_counter()
    tt = core.TrackerThread(dbp, 1, F, scene, mesh, False, False, bo)
    pump(tt, lambda m: results.append(m) if isinstance(m, core.FrameTrackingResult) else errors.append(m.what()))
    tt.join()
    dt = time.perf_counter() - t0
    t_err = max(float(np.abs(np.array(r.pose.t) - ts[r.frame - 1]).max()) for r in results) if results else None
    out_lines.append(json.dumps({"metric": "TrackerThread (Track Sequence) wall clock at %s" % args.config,
                                 "frames_tracked": len(results), "wall_s": dt, "frames_per_s": len(results) / dt,
                                 "max_abs_translation_error": t_err, "errors": errors}))

    # ---- RefinerThread (Refine Sequence, refiner.cc:692-725) --------------------------------------
    if results and len(results) == F - 1 and not args.no_refine:
        traj = core.CameraTrajectory(1, F)
        for k in range(1, F + 1):
            pz = core.Pose()
            if k in (1, F):
                cs = capi.camera_state(K, Rs[k - 1], ts[k - 1])
                pz.q, pz.t = [float(v) for v in cs.q], [float(v) for v in cs.t]
            else:
                r = results[k - 2]
                pz.q, pz.t = r.pose.q, r.pose.t
            traj.set(k, core.CameraState(intr, pz))
        updates = []
        t0 = time.perf_counter()
        rt = core.RefinerThread(dbp, traj, np.eye(4, dtype=np.float32), mesh, False, False, bo)
        pump(rt, lambda m: updates.append(m) if isinstance(m, core.RefineTrajectoryUpdate) else errors.append(m.what()))
        rt.join()
        dt = time.perf_counter() - t0
        out_lines.append(json.dumps({"metric": "RefinerThread (Refine Sequence) wall clock at %s" % args.config,
                                     "frames": F, "wall_s": dt, "updates": len(updates),
                                     "initial_cost": updates[-1].stats.initial_cost if updates else None,
                                     "final_cost": updates[-1].stats.cost if updates else None,
                                     "iterations": updates[-1].stats.iterations if updates else None,
                                     "errors": errors}))
    for l in out_lines:
        print(l)


if __name__ == "__main__":
    main()
