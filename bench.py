#!/usr/bin/env python
"""Headline benchmark: directed frame-pairs/s of the Analyze(+Track) hot path at 4K.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 4k|1080p|720p]

Contract (one JSON line on stdout, printed by rank 0): see the task description /
DESIGN.md section "Measurement".  A *step* is one pass of the hot path over a batch of
`--frames-per-step` consecutive synthetic frames (default 32 -> 256 directed pairs per step
in steady state): RGB->gray, 4-level pyramid, grid-thresholded Shi-Tomasi detection with
max_corners, pyramidal LK to the +-{1,2,4,8} neighbours, status filter.

  value : pairs/s with the RGB frames already resident in HBM (results stay on device).
  e2e   : the same through the reference-facing C ABI with HOST (pinned) frame buffers --
          the host->device copy of every frame and the device->host read of every
          keypoint/flow row are inside the timed region.
  roofline / cpu_baseline : see DESIGN.md.

`--impl reference` times the reference's own CPU path (OpenCV via cv2 + the restated
Polychase logic, oracle/analyze.py -- the C++ binary cannot be built in this image) on the
host cores, same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (width, height, max_corners, clip frames)
    "720p": (1280, 720, 2000, 64),
    "1080p": (1920, 1080, 4000, 1000),
    "4k": (3840, 2160, 8000, 1000),
}
SKIPS = (1, 2, 4, 8)


def pairs_for_new_frame(idx_in_clip: int) -> int:
    """Directed pairs that become computable when frame `idx_in_clip` (0-based) arrives."""
    return 2 * sum(1 for d in SKIPS if idx_in_clip - d >= 0)


def lk_algorithmic_bytes(w: int, h: int, n: int, n_out: int, win: int = 10, levels: int = 4) -> int:
    """per pair, for each of the 2 images and each level
    min(N*(win+2)^2, w_L*h_L) window bytes, plus 16 B per output row."""
    tot = 0
    lw, lh = w, h
    for _ in range(levels):
        tot += 2 * min(n * (win + 2) ** 2, lw * lh)
        lw, lh = (lw + 1) // 2, (lh + 1) // 2
    return tot + 16 * n_out


def frame_algorithmic_bytes(w: int, h: int, n: int, levels: int = 4) -> int:
    """Per frame: read RGB, write the gray pyramid, write keypoints."""
    tot = 3 * w * h + 8 * n
    lw, lh = w, h
    for _ in range(levels):
        tot += lw * lh
        lw, lh = (lw + 1) // 2, (lh + 1) // 2
    return tot


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []          # (sm_mhz, reasons bitmask) from NVML
        self.stop_flag = False

    # NVML in-process (a sample every few milliseconds: the timed region is a fraction of a second);
    # the nvidia-smi loop of the profiling recipe is the fallback
    NVML_REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def _nvml_loop(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((float(sm), int(rs)))
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1)
            sm = [x[0] for x in self.samples]
            reasons = set()
            for _, rs in self.samples:
                for bit, nm in self.NVML_REASONS.items():
                    if rs & bit:
                        reasons.add(nm)
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max,
                    "reasons": sorted(reasons), "samples": len(sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------
def read_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
def cpu_reference_run(cfg_name: str, steps: int, warmup: int, frames_per_step: int, threads: int):
    """The reference's CPU analyze loop (oracle/analyze.py) on host cores.  Each step analyzes
    `frames_per_step` consecutive frames of a resident window (all 8 neighbours present)."""
    import cv2
    from concurrent.futures import ThreadPoolExecutor
    from oracle import analyze as oanalyze, cvref, synth
    w, h, max_corners, _ = CONFIGS[cfg_name]
    cores = os.cpu_count() or 1
    cvref.pin(cores)
    total_frames = (steps + warmup) * frames_per_step
    window = 17 + frames_per_step        # frames kept so every analysed frame has its 8 partners
    clip = synth.Clip(w, h, window, seed=0)
    frames = {k: clip.rgb(k) for k in range(window)}
    gftt_kw = dict(max_corners=max_corners)
    pool = ThreadPoolExecutor(max_workers=threads)

    def run_step():
        pairs = 0
        for f in range(8, 8 + frames_per_step):
            _, rows = oanalyze.analyze_frame(lambda k: frames[k], f, 0, window, pool, gftt_kw, {})
            pairs += len(rows)
        return pairs

    for _ in range(warmup):
        run_step()
    t0 = time.perf_counter()
    pairs = 0
    for _ in range(steps):
        pairs += run_step()
    dt = time.perf_counter() - t0
    pool.shutdown()
    return pairs / dt, dt, pairs, cores, (f"{steps} steps x {frames_per_step} interior {cfg_name} frames "
                                           f"({pairs} directed pairs), cv2 {cv2.__version__} threads={cores}, "
                                           f"{threads} pairs in flight")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="4k", choices=sorted(CONFIGS))
    ap.add_argument("--frames-per-step", type=int, default=32)
    ap.add_argument("--cpu-baseline-frames", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-track", action="store_true", help="time detect+LK only (no per-frame PnP sweep)")
    ap.add_argument("--depth", type=int, default=16, help="frames in flight in the streaming analyzer")
    ap.add_argument("--diag", action="store_true",
                    help="also time upload-only and download-only legs and the raw H2D copy rate (stderr)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w, h, max_corners, clip_frames = CONFIGS[args.config]
    fps = args.frames_per_step
    track = not args.no_track
    stages = "detect+LK+PnP" if track else "detect+LK"
    metric = "frame-pairs/s (%s) at %s" % (stages, {"4k": "4K", "1080p": "1080p", "720p": "720p"}[args.config])
    config = {"workload": f"{args.config} synthetic clip, {max_corners} features/frame, detect+pyramid+LK "
                          f"(+-1,2,4,8 neighbours)" + (" + forward PnP sweep (ray cast + LM per frame)" if track else "")
                          + f", {fps} frames/step",
              "width": w, "height": h, "max_corners": max_corners, "frames_per_step": fps,
              "parallelism": f"frames sharded x{world}" if world > 1 else "single GPU",
              "frames_in_flight": args.depth,
              "host": "one process per GPU, bound to the GPU's CPU affinity (NVML)",
              "l2_policy": "inputs larger than L2 (each step streams fresh frames)"}

    if args.impl == "reference":
        if rank != 0:
            return
        ref_fps = min(fps, 2)           # bounded sample: a step is 2 interior frames (16 pairs)
        val, dt, pairs, cores, sample = cpu_reference_run(args.config, args.steps, max(args.warmup, 1), ref_fps, 4)
        line = {"metric": metric, "value": val, "unit": "frame-pairs/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32/f32", "data": "synthetic",
                "impl": "reference", "config": dict(config, frames_per_step=ref_fps),
                "cpu_baseline": {"value": val, "unit": "frame-pairs/s", "cores": cores, "kind": "port",
                                 "sample": sample},
                "e2e": {"value": val, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from polychase_b200 import synth   # input generator (texture + camera path)
    from polychase_b200 import capi

    torch.cuda.set_device(local_rank)
    orig_affinity = os.sched_getaffinity(0)
    try:       # run this rank (and allocate its pinned frame ring) on the CPUs next to its GPU
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    total_steps = args.steps + args.warmup
    # value leg: frames resident in HBM.  8 halo frames + all step frames, distinct frames.
    n_frames = 8 + total_steps * fps
    n_frames = min(n_frames, clip_frames + 8)
    ctx = capi.Context(device=local_rank, max_width=w, max_height=h, max_features=max(max_corners, 1024),
                       pipeline_depth=args.depth)
    tex = synth.make_texture(w, h, seed=0)
    ctx.synth_set_texture(tex)
    first = rank * clip_frames          # this rank's contiguous sub-sequence of the long clip
    K = synth.intrinsics(w, h)
    scale = synth.plane_scale(w, 4.0)
    Rs, ts = synth.camera_path(n_frames, 4.0, first - 8)
    stride = w * 3
    frame_bytes = stride * h
    dev_frames = ctx.device_alloc(frame_bytes * n_frames)
    for i in range(n_frames):
        ctx.synth_render(synth.homography(K, Rs[i], ts[i], w, h, scale), dev_frames + i * frame_bytes, stride)
    ctx.synchronize()

    gftt = capi.default_gftt(max_corners=max_corners)
    flow = capi.default_flow()

    def image_of(idx: int, ring: int) -> int:
        """Frame index -> resident image: a triangle wave over the ring, so consecutive frames are
        always consecutive images of the camera path (no motion jump where the ring wraps)."""
        if idx < ring:
            return idx
        period = 2 * (ring - 1)
        m = idx % period
        return m if m < ring else period - m

    # ---- Track (tracker.cc:36-213): the forward PnP sweep, chained on the device behind the analyzer
    # (pc_analyze_track_begin): each frame's ray cast + LM solve is queued right after its LK batch
    # and reads the flow rows / source poses where they already are in HBM.
    bundle = capi.default_bundle(loss_type=2)          # Cauchy: what the addon passes (blender_addon/operators/tracking.py:209)
    model = np.eye(4, dtype=np.float32)
    if track:
        verts, tris = synth.plane_mesh(w, h, scale)
        ctx.mesh_set(verts, tris)

    class Sweep:
        """Accumulates the tracked poses' statistics against the synthetic ground truth."""

        def __init__(self, ring: int):
            self.ring = ring
            self.tracked = 0
            self.matches = 0
            self.iterations = 0
            self.max_t_err = 0.0
            self.trace = []

        def truth(self, idx: int):
            i = image_of(idx, self.ring)
            return capi.camera_state(K, Rs[i], ts[i])

        def consume(self, r, timed: bool):
            if r["tracked"] != 1 or not timed:
                return
            idx = r["frame_id"] - (first - 8)
            self.tracked += 1
            self.matches += r["num_matches"]
            self.iterations += r["stats"].iterations
            i = image_of(idx, self.ring)
            e = float(np.abs(np.array(r["camera"].t[:]) - ts[i]).max())
            self.max_t_err = max(self.max_t_err, e)
            if args.diag:
                self.trace.append((idx, round(e, 5), int(r["stats"].iterations), int(r["num_matches"]),
                                   round(float(np.abs(ts[i]).max()), 3)))

    def run_steps(n_steps: int, start_frame_idx: int, mem_kind: int, base_ptr: int, ring: int, download: bool,
                  sweep=None, timed=False):
        """Pushes n_steps*fps frames (cycling over `ring` resident buffers) and pops results."""
        pairs = 0
        rows = 0
        idx = start_frame_idx
        for _ in range(n_steps * fps):
            ctx.analyze_push(first - 8 + idx, base_ptr + image_of(idx, ring) * frame_bytes, stride, mem_kind)
            idx += 1
            if ctx.analyze_pending() >= args.depth:
                r = ctx.analyze_pop(download=download, copy=False)
                pairs += len(r["pairs"])
                rows += sum(p[2] for p in r["pairs"])
                if sweep is not None:
                    sweep.consume(r, timed)
        return pairs, rows, idx

    def drain(download: bool, sweep=None, timed=False):
        pairs = rows = 0
        while ctx.analyze_pending():
            r = ctx.analyze_pop(download=download, copy=False)
            pairs += len(r["pairs"])
            rows += sum(p[2] for p in r["pairs"])
            if sweep is not None:
                sweep.consume(r, timed)
        return pairs, rows

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_leg(mem_kind: int, base_ptr: int, ring: int, download: bool):
        sweep = Sweep(ring) if track else None
        ctx.analyze_begin(w, h, first - 8, 10 ** 6, gftt, flow)
        ctx.analyze_set_halo(8)
        if track:                                            # the sweep starts from the shard's known pose:
            ctx.analyze_track_begin(model, bundle)           # the frame before its first own frame
            ctx.analyze_track_seed(first - 1, sweep.truth(7))
        idx = 0
        for _ in range(8):                                   # halo frames of the previous shard
            ctx.analyze_push(first - 8 + idx, base_ptr + image_of(idx, ring) * frame_bytes, stride, mem_kind)
            idx += 1
            if ctx.analyze_pending() >= args.depth:
                r = ctx.analyze_pop(download=download, copy=False)
                if sweep is not None:
                    sweep.consume(r, False)
        _, _, idx = run_steps(args.warmup, idx, mem_kind, base_ptr, ring, download, sweep)
        drain(download, sweep)
        ctx.timing_read(reset=True)
        ctx.timing_enable(True)
        launches0 = ctx.kernel_launches()
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        ctx.mark(0)
        t0 = time.perf_counter()
        pairs, rows, idx = run_steps(args.steps, idx, mem_kind, base_ptr, ring, download, sweep, True)
        p2, r2 = drain(download, sweep, True)                # pops wait for each frame's rows and pose,
        ctx.mark(1)                                          # so mark 1 is recorded after all the work
        ctx.synchronize()
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop()
        dev_ms = ctx.elapsed_ms(0, 1)
        times = ctx.timing_read(reset=True)
        ctx.timing_enable(False)
        launches = ctx.kernel_launches() - launches0
        ctx.analyze_end()
        out = dict(pairs=pairs + p2, rows=rows + r2, dev_ms=dev_ms, wall_s=wall, clocks=clocks, times=times,
                   launches=launches)
        if sweep is not None:
            if args.diag and rank == 0:
                import sys
                print(json.dumps({"track_trace(idx, |t err|, LM iters, matches, |t|max)": sweep.trace[::16]}), file=sys.stderr)
            out["track"] = {"frames_tracked": sweep.tracked,
                            "matches_per_frame": sweep.matches / max(sweep.tracked, 1),
                            "lm_iterations_per_frame": sweep.iterations / max(sweep.tracked, 1),
                            "max_abs_translation_error": sweep.max_t_err}
        return out

    # ---- value: inputs resident in HBM, results stay on device --------------------------
    res = timed_leg(capi.PC_MEM_DEVICE, dev_frames, n_frames, download=False)

    # ---- e2e: pinned host frames in, rows out -------------------------------------------
    e2e = None
    if not args.no_e2e:
        ring = 32
        host_ptr = ctx.pinned_alloc(frame_bytes * ring)
        import ctypes
        for i in range(ring):                                # fill the pinned ring from the device clip
            ctx.lib.pc_memcpy_d2h(ctx.h, ctypes.c_void_p(host_ptr + i * frame_bytes),
                                  ctypes.c_void_p(dev_frames + i * frame_bytes), frame_bytes)
        e2e_res = timed_leg(capi.PC_MEM_HOST_PINNED, host_ptr, ring, download=True)
        if args.diag and rank == 0:
            import sys
            up = timed_leg(capi.PC_MEM_HOST_PINNED, host_ptr, ring, download=False)
            down = timed_leg(capi.PC_MEM_DEVICE, dev_frames, n_frames, download=True)
            src = torch.empty(frame_bytes, dtype=torch.uint8).pin_memory()
            dst = torch.empty(frame_bytes, dtype=torch.uint8, device="cuda")
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            ev[0].record()
            for _ in range(20):
                dst.copy_(src, non_blocking=True)
            ev[1].record()
            torch.cuda.synchronize()
            h2d_gbs = 20 * frame_bytes / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9
            print(json.dumps({"diag": {"resident_ms_per_step": res["dev_ms"] / args.steps,
                                       "resident_wall_ms_per_step": 1e3 * res["wall_s"] / args.steps,
                                       "e2e_ms_per_step": e2e_res["dev_ms"] / args.steps,
                                       "upload_only_ms_per_step": up["dev_ms"] / args.steps,
                                       "download_only_ms_per_step": down["dev_ms"] / args.steps,
                                       "h2d_pinned_gbs": h2d_gbs,
                                       "h2d_ms_per_step_at_that_rate": fps * frame_bytes / h2d_gbs / 1e6}}),
                  file=sys.stderr)
        ctx.pinned_free(host_ptr)
        e2e = e2e_res

    def reduce_max(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    dev_s = reduce_max(res["dev_ms"]) * 1e-3
    total_pairs = reduce_sum(res["pairs"])
    value = total_pairs / dev_s
    line = {"metric": metric, "value": value, "unit": "frame-pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32/f32", "data": "synthetic",
            "config": config, "clocks": res["clocks"], "gpu_launches": int(res["launches"]),
            "wall_s": res["wall_s"], "pairs": int(total_pairs)}
    if "track" in res:
        line["track"] = res["track"]
    if e2e is not None:
        e_s = reduce_max(e2e["dev_ms"]) * 1e-3
        e_pairs = reduce_sum(e2e["pairs"])
        rows_per_step = e2e["rows"] / args.steps
        line["e2e"] = {"value": e_pairs / e_s, "unit": "frame-pairs/s",
                       "h2d_bytes_per_step": int(frame_bytes * fps),
                       "d2h_bytes_per_step": int(rows_per_step * 16 + fps * max_corners * 8),
                       "wall_s": e2e["wall_s"], "clocks": e2e["clocks"]}
        if "track" in e2e:
            line["e2e"]["track"] = e2e["track"]

    # ---- roofline of the dominant kernel (rank 0's numbers) ------------------------------
    peak, peak_src = read_peak_hbm()
    t = res["times"]
    fam = {k[:-3]: (t[k], t[k[:-3] + "_n"]) for k in t if k.endswith("_ms")}
    # dominant = most SM-time: the PnP solve is one 8-CTA cluster (8 of the SMs) on its own stream,
    # every other family fills the chip while it runs
    sm_share = {"pnp": 8.0 / 148.0}
    dominant = max(fam, key=lambda k: fam[k][0] * sm_share.get(k, 1.0))
    step_ms = sum(v[0] for v in fam.values())
    n_out = res["rows"] / max(res["pairs"], 1)
    per_launch = {
        "lk": 8 * lk_algorithmic_bytes(w, h, max_corners, int(n_out)),
        "gray_pyr": frame_algorithmic_bytes(w, h, 0),
        "min_eig": w * h + 4 * w * h,                 # read gray, write the eig map (materialised)
        "select": 4 * w * h + 8 * max_corners,
        "compact": 8 * max_corners * 13 + 8 * int(n_out) * 16,
    }
    if "track" in res:                                    # 20 B per match, read once per pass
        m_frame = res["track"]["matches_per_frame"]
        per_launch["raycast"] = int(20 * m_frame)
        per_launch["pnp"] = int(20 * m_frame)
    dom_ms, dom_n = fam[dominant]
    avg_ms = dom_ms / max(dom_n, 1)
    achieved = per_launch.get(dominant, 0) / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    traffic = None                                        # measured DRAM bytes per launch (ncu --set full), if captured
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(dominant)
        traffic = traffic and (traffic["bytes_per_launch"] if args.config == "4k" else None)
    except Exception:
        traffic = None
    line["roofline"] = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                        "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": per_launch.get(dominant, 0),
                        "share_of_step": dom_ms / step_ms if step_ms else None,
                        "per_kernel": {k: {"ms_total": v[0], "launch_groups": int(v[1]),
                                           "avg_ms": v[0] / max(v[1], 1),
                                           "achieved_gbs": (per_launch.get(k, 0) / (v[0] / max(v[1], 1) * 1e-3) / 1e9)
                                           if v[0] > 0 else 0.0}
                                       for k, v in fam.items() if v[1]}}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) -----------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:                                              # the CPU arm gets every host core back
            os.sched_setaffinity(0, orig_affinity)
        except Exception:
            pass
        val, dt, pairs, cores, sample = cpu_reference_run(args.config, 1, 1, args.cpu_baseline_frames, 4)
        line["cpu_baseline"] = {"value": val, "unit": "frame-pairs/s", "cores": cores, "kind": "port",
                                "sample": sample}
    if world > 1:
        # the only collective of the design: stitch per-GPU trajectory segments (64 B/frame)
        seg = torch.zeros((clip_frames, 16), device="cuda", dtype=torch.float32)
        out = [torch.empty_like(seg) for _ in range(world)]
        dist.all_gather(out, seg)
        torch.cuda.synchronize()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))
    ctx.device_free(dev_frames)
    ctx.close()


if __name__ == "__main__":
    main()
