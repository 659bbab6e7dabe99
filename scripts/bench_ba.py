"""Refine at BASELINE.json configs[4] scale: a 200-frame 4K segment, 8000 keypoints/frame,
8*200-30 = 1570 directed edges (~12.5 M residual rows), Cauchy(1.0), intrinsics fixed.

Analyze (detect + LK) runs on the GPU to produce the flows -- that is the reference's data path
(Analyze writes the DB, Refine reads it through CachedDatabase, refiner.cc:97-161) -- then the
trajectory (ground truth + N(0, 0.2 deg / 0.5 % depth) on interior frames, seed 1) is refined with
pc_ba_solve.  Prints one JSON line: cost evaluations/s, normal-equation builds/s, residual rows/s,
iterations, and the pose error against the synthetic ground truth before / after.

    python scripts/bench_ba.py [--frames 200] [--config 4k]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {"4k": (3840, 2160, 8000), "1080p": (1920, 1080, 4000), "720p": (1280, 720, 2000)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--config", default="4k", choices=sorted(CONFIGS))
    ap.add_argument("--max-iterations", type=int, default=20)
    ap.add_argument("--reps", type=int, default=5, help="timed repetitions of the cost / build entry points")
    ap.add_argument("--motion", default="survey", choices=["survey", "r1"],
                    help="survey: image motion <~ 20 px per 8 frames (SURVEY.md 8d, synth.survey_speed); "
                         "r1: round 1's path (up to 89 px per 8 frames at 4K, outside LK's capture range)")
    args = ap.parse_args()

    from polychase_b200 import synth  # input generator
    from polychase_b200 import capi
    from polychase_b200.geometry import quat_from_matrix

    w, h, max_corners = CONFIGS[args.config]
    F = args.frames
    ctx = capi.Context(device=0, max_width=w, max_height=h, max_features=max(max_corners, 1024), pipeline_depth=4)
    tex = synth.make_texture(w, h, seed=0)
    ctx.synth_set_texture(tex)
    K = synth.intrinsics(w, h)
    scale = synth.plane_scale(w, 4.0)
    speed = synth.survey_speed(w) if args.motion == "survey" else 1.0
    Rs, ts = synth.camera_path(F, 4.0, 0, speed)
    stride = w * 3
    frame_bytes = stride * h
    ring = 16
    dev = ctx.device_alloc(frame_bytes * ring)

    # ---- Analyze: flows of the whole segment (frames rendered just ahead of the analyzer) -------
    gftt = capi.default_gftt(max_corners=max_corners)
    flow = capi.default_flow()
    kps = {}
    flows = {}

    def take(r):
        kps[r["frame_id"]] = np.array(r["keypoints"], np.float32).reshape(-1, 2).copy()
        for (a, b, rows, idx, tgt, err) in r["pairs"]:
            flows[(a, b)] = (np.array(idx, np.uint32).copy(), np.array(tgt, np.float32).reshape(-1, 2).copy())

    t0 = time.perf_counter()
    ctx.analyze_begin(w, h, 0, F, gftt, flow)
    for i in range(F):
        slot = dev + (i % ring) * frame_bytes
        ctx.synth_render(synth.homography(K, Rs[i], ts[i], w, h, scale), slot, stride)
        ctx.synchronize()
        ctx.analyze_push(i, slot, stride, capi.PC_MEM_DEVICE)
        if ctx.analyze_pending() >= 4:
            take(ctx.analyze_pop(download=True, copy=True))
    while ctx.analyze_pending():
        take(ctx.analyze_pop(download=True, copy=True))
    ctx.analyze_end()
    analyze_s = time.perf_counter() - t0

    # ---- Refine problem (GlobalRefinementProblem, refiner.cc:561-690) ------------------------------
    verts, tris = synth.plane_mesh(w, h, scale)
    ctx.mesh_set(verts, tris)
    edges = [(a, b, flows[(a, b)][0], flows[(a, b)][1]) for (a, b) in sorted(flows) if len(flows[(a, b)][0])]
    rows = int(sum(len(e[2]) for e in edges))
    model = np.eye(4, dtype=np.float32)
    t0 = time.perf_counter()
    ctx.ba_load([kps[k] for k in range(F)], edges, model, False, False)
    load_s = time.perf_counter() - t0

    rng = np.random.default_rng(1)
    truth = [capi.camera_state(K, Rs[i], ts[i]) for i in range(F)]
    traj = []
    for i in range(F):
        R, t = Rs[i], ts[i]
        if 0 < i < F - 1:
            wv = rng.normal(0, np.deg2rad(0.2), 3)
            th = np.linalg.norm(wv)
            kx = np.array([[0, -wv[2], wv[1]], [wv[2], 0, -wv[0]], [-wv[1], wv[0], 0]])
            dR = np.eye(3) + (np.sin(th) / th) * kx + ((1 - np.cos(th)) / th ** 2) * (kx @ kx)
            R = R @ dR
            t = t + rng.normal(0, 0.005 * 4.0, 3)
        traj.append(capi.camera_state(K, R, t))

    def pose_err(tr):
        et = max(float(np.abs(np.array(tr[i].t[:]) - np.array(truth[i].t[:])).max()) for i in range(F))
        eq = 0.0
        for i in range(F):
            qa, qb = np.array(tr[i].q[:], np.float64), np.array(truth[i].q[:], np.float64)
            if np.dot(qa, qb) < 0:
                qb = -qb
            eq = max(eq, float(np.abs(qa - qb).max()))
        return et, eq

    bo = capi.default_bundle(loss_type=2, max_iterations=args.max_iterations)
    truth_cost = ctx.ba_cost(truth, bo)                       # what a perfect refine would reach
    ctx.ba_cost(traj, bo)                                     # warm-up (fills the primitive-id cache)
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        ctx.ba_cost(traj, bo)
    cost_s = (time.perf_counter() - t0) / args.reps
    ctx.ba_normal_equations(traj, bo)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        ctx.ba_normal_equations(traj, bo)
    build_s = (time.perf_counter() - t0) / args.reps

    ctx.timing_read(reset=True)
    ctx.timing_enable(True)
    e0 = pose_err(traj)
    t0 = time.perf_counter()
    out, st = ctx.ba_solve(traj, bo)
    solve_s = time.perf_counter() - t0
    times = ctx.timing_read(reset=True)
    ctx.timing_enable(False)
    e1 = pose_err(out)
    line = {
        "metric": "refine (bundle adjust) at %s" % args.config, "frames": F, "edges": len(edges), "residual_rows": rows,
        "keypoints_per_frame": max_corners, "loss": "cauchy(1.0)", "optimize_intrinsics": False,
        "analyze_wall_s": analyze_s, "ba_load_wall_s": load_s,
        "cost_eval_ms": 1e3 * cost_s, "cost_rows_per_s": rows / cost_s,
        "normal_equation_build_ms_incl_download": 1e3 * build_s, "builds_per_s": 1.0 / build_s,
        "build_rows_per_s": rows / build_s,
        "solve_wall_s": solve_s, "solve_gpu_kernel_ms": times.get("ba_ms"), "solve_kernel_spans": times.get("ba_n"),
        "iterations": int(st.iterations), "invalid_steps": int(st.invalid_steps),
        "initial_cost": float(st.initial_cost), "final_cost": float(st.cost), "cost_at_ground_truth": truth_cost,
        "lambda": float(st.lambda_), "grad_norm": float(st.grad_norm), "step_norm": float(st.step_norm),
        "max_abs_t_err_before": e0[0], "max_abs_t_err_after": e1[0],
        "max_abs_q_err_before": e0[1], "max_abs_q_err_after": e1[1],
        "motion": args.motion,
    }
    print(json.dumps(line))
    ctx.device_free(dev)
    ctx.close()


if __name__ == "__main__":
    main()
