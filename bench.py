#!/usr/bin/env python
"""Headline benchmark: directed frame-pairs/s of the Analyze(+Track) hot path at 4K.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 4k|1080p|720p]
                    [--motion survey|r1]

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
  ba        : (N=1) Refine at BASELINE configs[4] scale -- 200 keyframes x 8000 tracks at 4K: cost / normal-equation
              build / LM-iteration times, rows/s, HBM fraction of the build kernel, pose error against the truth.
  plugin_e2e: (N=1) the same Analyze pass through the reference's own entry point, polychase_core.OpticalFlowThread
              (frame request / provide_frame hand-off, SQLite database written), wall clock.
  collective: (N>1) the path's one collective -- all-gather of the poses the timed sweep produced
              (polychase_b200.shard.allgather_trajectory), timed on its own and verified against the per-rank results.

Synthetic clip: the camera path is scaled in time so that the image moves by at most
about 20 px over 8 frames at the configured width (`--motion survey`, polychase_b200/synth.py::survey_speed).
Round 1's clip (`--motion r1`) moved the 4K image by up to 89 px per 8 frames, outside the capture range of the
reference's 4-level / 10x10 LK: 96 % of its skip-8 tracks were wrong by > 3 px (profiles/r2_lk_track_error_by_skip.txt).
For N > 1 the clip is BASELINE configs[3]'s 4000 frames, split with polychase_b200.shard.shard_range; every rank
times its own halo (the previous shard's last 8 frames) inside the timed region.

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


def proportional_shares(rates, frames_per_step: int, floor: int = 8):
    """Frames per step of every rank, proportional to its measured frame rate: the whole job keeps
    len(rates) * frames_per_step frames per step (up to rounding); no rank gets fewer than `floor`."""
    total = float(sum(rates))
    n = len(rates)
    if total <= 0.0:
        return [frames_per_step] * n
    return [max(floor, int(round(frames_per_step * n * float(r) / total))) for r in rates]


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
CPU_ARM_ENV = {"MALLOC_MMAP_THRESHOLD_": str(32 << 20)}


def cpu_arm_env_ready() -> bool:
    """glibc serves blocks above its mmap threshold with a fresh mmap (page faults on every touch) and only
    raises that threshold as mmapped blocks get freed; depending on what the Python process allocated before,
    cv2's per-call image buffers (MBs each) stay in that slow mode for the whole bounded sample (measured: 2.7x
    at 720p).  A long-running reference process ends up with the threshold at its 32 MB maximum, so the CPU arm
    runs in a process started with that setting (favourable to the CPU arm: conservative for the GPU/CPU ratio)."""
    return all(os.environ.get(k) == v for k, v in CPU_ARM_ENV.items())


def cpu_reference_run(cfg_name: str, steps: int, warmup: int, frames_per_step: int, threads: int, speed: float = 1.0,
                      track: bool = True):
    """The reference's CPU path on host cores: the Analyze loop (oracle/analyze.py: cv2 + the restated Polychase
    logic, with the reference's redundancy) and, for "detect+LK+PnP", SolveFrame of every analysed frame
    (oracle/track_port.c: BVH ray cast of the matched keypoints of the posed sources f-1, f-2, f-4, f-8 + the
    float32 LM solve, tracker.cc:36-131).  Each step analyses `frames_per_step` consecutive frames of a resident
    window (all 8 neighbours present).  The flows INTO a frame that its Track step consumes were written by the
    Analyze iterations of its source frames, which lie before the bounded sample: they are produced untimed."""
    import cv2
    from concurrent.futures import ThreadPoolExecutor
    from oracle import analyze as oanalyze, cvref, synth
    w, h, max_corners, _ = CONFIGS[cfg_name]
    cores = os.cpu_count() or 1
    cvref.pin(cores)
    window = 17 + frames_per_step        # frames kept so every analysed frame has its 8 partners
    clip = synth.Clip(w, h, window, seed=0, speed=speed)
    frames = {k: clip.rgb(k) for k in range(window)}
    gftt_kw = dict(max_corners=max_corners)
    pool = ThreadPoolExecutor(max_workers=threads)
    track_in = {}
    if track:
        from oracle import geometry as G
        from oracle import pnp as opnp
        from oracle import track_port
        mesh = track_port.Mesh(clip.verts, clip.tris)
        bopts = opnp.BundleOptions(loss_type=opnp.CAUCHY)          # what the addon passes (operators/tracking.py:209)
        model = np.eye(4, dtype=np.float32)

        def cam_of(k):
            it = G.Intrinsics(clip.K["fx"], clip.K["fy"], clip.K["cx"], clip.K["cy"], 1.0, w, h, G.OPENCV).f32()
            return G.CameraState(it, G.Pose(G.quat_from_matrix(clip.R[k]).astype(np.float32), clip.t[k].astype(np.float32)))

        grays, kps_of = {}, {}
        for f in range(8, 8 + frames_per_step):                    # untimed: the rows the Track step of frame f reads
            srcs = []
            for d in (8, 4, 2, 1):
                a = f - d
                if a not in grays:
                    grays[a] = cvref.rgb2gray(frames[a])
                if a not in kps_of:
                    kps_of[a], _ = cvref.gftt(grays[a], **gftt_kw)
                if f not in grays:
                    grays[f] = cvref.rgb2gray(frames[f])
                nxt, st, _ = cvref.lk(grays[a], grays[f], kps_of[a])
                ok = st == 1
                srcs.append((cam_of(a), kps_of[a], np.nonzero(ok)[0].astype(np.uint32), nxt[ok]))
            track_in[f] = srcs

    def run_step():
        pairs = 0
        for f in range(8, 8 + frames_per_step):
            _, rows = oanalyze.analyze_frame(lambda k: frames[k], f, 0, window, pool, gftt_kw, {})
            pairs += len(rows)
            if track:
                r = track_port.track_frame(mesh, model, track_in[f], cam_of(f - 1), bopts)
                assert r is not None and r[2] > 0.5
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
                                           f"{threads} pairs in flight"
                                           + ("; Track: SolveFrame per frame, plain-C port, 1 thread" if track else ""))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="4k", choices=sorted(CONFIGS))
    ap.add_argument("--motion", default="survey", choices=["survey", "r1"],
                    help="survey: image motion <~ 20 px per 8 frames (SURVEY.md 8d); r1: round 1's clip")
    ap.add_argument("--frames-per-step", type=int, default=32)
    ap.add_argument("--cpu-baseline-frames", type=int, default=4)
    ap.add_argument("--ref-frames-per-step", type=int, default=2, help="frames per step of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-track", action="store_true", help="time detect+LK only (no per-frame PnP sweep)")
    ap.add_argument("--no-ba", action="store_true", help="skip the configs[4] refine block")
    ap.add_argument("--no-plugin", action="store_true", help="skip the polychase_core.OpticalFlowThread block")
    ap.add_argument("--ba-frames", type=int, default=200)
    ap.add_argument("--plugin-frames", type=int, default=1000)
    ap.add_argument("--plugin-only", action="store_true", help="run only the OpticalFlowThread block (used by the main run)")
    ap.add_argument("--depth", type=int, default=16, help="frames in flight in the streaming analyzer")
    ap.add_argument("--equal-shards", action="store_true",
                    help="N > 1, e2e leg: 32 frames per rank and step instead of shares proportional to the upload rates")
    ap.add_argument("--diag", action="store_true",
                    help="also time upload-only and download-only legs and the raw H2D copy rate (stderr)")
    args = ap.parse_args()

    if args.plugin_only:
        plugin_only(args)
        return
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w, h, max_corners, clip_frames = CONFIGS[args.config]
    if world > 1:
        clip_frames *= 4                 # BASELINE configs[3]: the 4000-frame clip, sharded
    fps = args.frames_per_step
    track = not args.no_track
    stages = "detect+LK+PnP" if track else "detect+LK"
    metric = "frame-pairs/s (%s) at %s" % (stages, {"4k": "4K", "1080p": "1080p", "720p": "720p"}[args.config])
    from polychase_b200 import synth   # input generator (texture + camera path)
    speed = synth.survey_speed(w) if args.motion == "survey" else 1.0
    config = {"workload": f"{args.config} synthetic {clip_frames}-frame clip, {max_corners} features/frame, detect+pyramid+LK "
                          f"(+-1,2,4,8 neighbours)" + (" + forward PnP sweep (ray cast + LM per frame)" if track else "")
                          + f", {fps} frames/step",
              "width": w, "height": h, "max_corners": max_corners, "frames_per_step": fps,
              "motion": ("image motion <= ~20 px per 8 frames (SURVEY.md 8d; camera path time scale %.3f)" % speed)
                        if args.motion == "survey" else "round-1 clip (up to 89 px per 8 frames at 4K)",
              "parallelism": (f"frames sharded x{world} (polychase_b200.shard.shard_range over {clip_frames} frames, "
                              f"8-frame halo inside the timed region)") if world > 1 else "single GPU",
              "frames_in_flight": args.depth,
              "host": "one process per GPU, bound to the GPU's CPU affinity (NVML)",
              "l2_policy": "inputs larger than L2 (each step streams fresh frames)"}

    if args.impl == "reference":
        if rank != 0:
            return
        if not cpu_arm_env_ready():     # see cpu_arm_env_ready: re-exec with the allocator in its steady state
            os.execve(sys.executable, [sys.executable, os.path.abspath(__file__)] + sys.argv[1:],
                      dict(os.environ, **CPU_ARM_ENV))
        ref_fps = min(fps, args.ref_frames_per_step)     # bounded sample: a step is 2 interior frames (16 pairs)
        val, dt, pairs, cores, sample = cpu_reference_run(args.config, args.steps, max(args.warmup, 1), ref_fps, 4,
                                                          speed, track)
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
    from polychase_b200 import capi
    from polychase_b200 import shard

    torch.cuda.set_device(local_rank)
    orig_affinity = os.sched_getaffinity(0)
    numa = {}
    try:       # run this rank (and allocate its pinned frame ring) on the CPUs next to its GPU
        import pynvml
        pynvml.nvmlInit()
        hdl = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(hdl)
        numa["cpus"] = len(os.sched_getaffinity(0))
        try:
            numa["gpu_numa_node"] = int(open("/sys/bus/pci/devices/%s/numa_node" %
                                             pynvml.nvmlDeviceGetPciInfo(hdl).busId.lower().replace("00000000:", "0000:")).read())
        except Exception:
            pass
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    total_steps = args.steps + args.warmup
    start, count = shard.shard_range(0, clip_frames, world, rank)      # this rank's contiguous sub-sequence
    halo = shard.halo_frames(0, start)                                # the previous shard's last frames (0 for rank 0)
    # resident images: the halo + every frame of the longer pass, capped at the shard (then a triangle wave)
    n_frames = halo + min(max(total_steps * fps, args.ba_frames if world == 1 else 0), count)
    ctx = capi.Context(device=local_rank, max_width=w, max_height=h, max_features=max(max_corners, 1024),
                       pipeline_depth=args.depth)
    tex = synth.make_texture(w, h, seed=0)
    ctx.synth_set_texture(tex)
    K = synth.intrinsics(w, h)
    scale = synth.plane_scale(w, 4.0)
    Rs, ts = synth.camera_path(n_frames, 4.0, start - halo, speed)
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
    verts, tris = synth.plane_mesh(w, h, scale)
    if track:
        ctx.mesh_set(verts, tris)

    class Sweep:
        """Accumulates the tracked poses and their statistics against the synthetic ground truth."""

        def __init__(self, ring: int):
            self.ring = ring
            self.tracked = 0
            self.matches = 0
            self.iterations = 0
            self.max_t_err = 0.0
            self.trace = []
            self.poses = {}                      # frame id -> 16 floats (pc_camera_state), timed frames only

        def truth(self, idx: int):
            i = image_of(idx, self.ring)
            return capi.camera_state(K, Rs[i], ts[i])

        def consume(self, r, timed: bool):
            if not r["tracked"] or not timed:
                return
            self.poses[r["frame_id"]] = np.frombuffer(bytes(r["camera"]), np.float32).copy()
            if r["tracked"] != 1:
                return
            idx = r["frame_id"] - (start - halo)
            self.tracked += 1
            self.matches += r["num_matches"]
            self.iterations += r["stats"].iterations
            i = image_of(idx, self.ring)
            e = float(np.abs(np.array(r["camera"].t[:]) - ts[i]).max())
            self.max_t_err = max(self.max_t_err, e)
            if args.diag:
                self.trace.append((idx, round(e, 5), int(r["stats"].iterations), int(r["num_matches"]),
                                   round(float(np.abs(ts[i]).max()), 3)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass(n_steps: int, mem_kind: int, base_ptr: int, ring: int, download: bool, sweep, timed: bool, fps: int = fps):
        """One Analyze(+Track) pass over this rank's sub-sequence: the halo, then n_steps * fps own frames."""
        ctx.analyze_begin(w, h, start - halo, 10 ** 6, gftt, flow)
        ctx.analyze_set_halo(halo)
        if track:                                            # the sweep starts from the shard's known pose: the frame
            ctx.analyze_track_begin(model, bundle)           # before its first own frame (rank 0: the clip's first frame)
            if halo:
                ctx.analyze_track_seed(start - 1, sweep.truth(halo - 1))
            else:
                ctx.analyze_track_seed(start, sweep.truth(0))
        pairs = rows = 0

        def take():
            nonlocal pairs, rows
            r = ctx.analyze_pop(download=download, copy=False)
            pairs += len(r["pairs"])
            rows += sum(p[2] for p in r["pairs"])
            if sweep is not None:
                sweep.consume(r, timed)

        for idx in range(halo + n_steps * fps):
            ctx.analyze_push(start - halo + idx, base_ptr + image_of(idx, ring) * frame_bytes, stride, mem_kind)
            if ctx.analyze_pending() >= args.depth:
                take()
        while ctx.analyze_pending():                         # pops wait for each frame's rows and pose
            take()
        return pairs, rows

    def timed_leg(mem_kind: int, base_ptr: int, ring: int, download: bool, fps: int = fps):
        sweep = Sweep(ring) if track else None
        one_pass(args.warmup, mem_kind, base_ptr, ring, download, sweep, False, fps)      # W untimed warm-up steps
        ctx.analyze_end()
        ctx.timing_read(reset=True)
        ctx.timing_enable(True)
        launches0 = ctx.kernel_launches()
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        ctx.mark(0)
        t0 = time.perf_counter()
        pairs, rows = one_pass(args.steps, mem_kind, base_ptr, ring, download, sweep, True, fps)   # K steps (+ the halo)
        ctx.mark(1)                                          # recorded after all the work
        ctx.synchronize()
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop()
        dev_ms = ctx.elapsed_ms(0, 1)
        times = ctx.timing_read(reset=True)
        ctx.timing_enable(False)
        launches = ctx.kernel_launches() - launches0
        ctx.analyze_end()
        out = dict(pairs=pairs, rows=rows, dev_ms=dev_ms, wall_s=wall, clocks=clocks, times=times, launches=launches)
        if sweep is not None:
            if args.diag and rank == 0:
                print(json.dumps({"track_trace(idx, |t err|, LM iters, matches, |t|max)": sweep.trace[::16]}), file=sys.stderr)
            out["track"] = {"frames_tracked": sweep.tracked,
                            "matches_per_frame": sweep.matches / max(sweep.tracked, 1),
                            "lm_iterations_per_frame": sweep.iterations / max(sweep.tracked, 1),
                            "max_abs_translation_error": sweep.max_t_err}
            out["poses"] = sweep.poses
        return out

    # ---- value: inputs resident in HBM, results stay on device --------------------------
    res = timed_leg(capi.PC_MEM_DEVICE, dev_frames, n_frames, download=False)

    # ---- e2e: pinned host frames in, rows out -------------------------------------------
    e2e = None
    h2d_gbs = None
    if not args.no_e2e:
        ring = min(32, n_frames)
        host_ptr = ctx.pinned_alloc(frame_bytes * ring)
        import ctypes
        for i in range(ring):                                # fill the pinned ring from the device clip
            ctx.lib.pc_memcpy_d2h(ctx.h, ctypes.c_void_p(host_ptr + i * frame_bytes),
                                  ctypes.c_void_p(dev_frames + i * frame_bytes), frame_bytes)
        # this rank's raw pinned host->device copy rate, all ranks copying at once (what bounds e2e)
        src = torch.empty(frame_bytes, dtype=torch.uint8).pin_memory()
        dst = torch.empty(frame_bytes, dtype=torch.uint8, device="cuda")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        dst.copy_(src, non_blocking=True)
        barrier()
        ev[0].record()
        for _ in range(20):
            dst.copy_(src, non_blocking=True)
        ev[1].record()
        torch.cuda.synchronize()
        h2d_gbs = 20 * frame_bytes / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9
        del src, dst
        # Ingest-proportional shards (N > 1): the GPUs of one box do not see the same host link -- with eight ranks
        # uploading, four B200s of the pool's boxes receive 23-25 GB/s and four 35-39 GB/s (profiles/r2_n_h2d_numa_probe.json)
        # -- and with equal shards the slow ranks set the pace.  Every rank takes a share of the step's
        # world * frames_per_step frames proportional to the frame rate it reaches in two short untimed passes with all
        # ranks running (equal shares first, then the shares that gave), so all ranks finish together; the work of
        # the whole job per step is unchanged.  --equal-shards keeps 32 frames per rank.
        e2e_fps = fps
        e2e_fps_all = [fps] * world
        if world > 1 and not args.equal_shards:
            def measured_rates(fps_r: int):
                """frames/s of every rank for one untimed 2-step e2e pass with fps_r frames per step on this rank,
                all ranks running at once (what the shares are proportional to)."""
                barrier()
                ctx.mark(2)
                one_pass(2, capi.PC_MEM_HOST_PINNED, host_ptr, ring, True, Sweep(ring) if track else None, False, fps_r)
                ctx.mark(3)
                ctx.synchronize()
                ctx.analyze_end()
                r = torch.zeros(world, dtype=torch.float64, device="cuda")
                r[rank] = (halo + 2 * fps_r) / max(ctx.elapsed_ms(2, 3), 1e-3)
                dist.all_reduce(r)
                return r.cpu().numpy()

            def shares(rates):
                return proportional_shares([float(x) for x in rates], fps)

            one_pass(1, capi.PC_MEM_HOST_PINNED, host_ptr, ring, True, Sweep(ring) if track else None, False, fps)   # first-touch costs
            ctx.analyze_end()
            e2e_fps_all = shares(measured_rates(fps))                       # from equal shards ...
            e2e_fps_all = shares(measured_rates(e2e_fps_all[rank]))         # ... refined once under the new load
            e2e_fps = e2e_fps_all[rank]
        e2e_res = timed_leg(capi.PC_MEM_HOST_PINNED, host_ptr, ring, download=True, fps=e2e_fps)
        if args.diag and rank == 0:
            up = timed_leg(capi.PC_MEM_HOST_PINNED, host_ptr, ring, download=False)
            down = timed_leg(capi.PC_MEM_DEVICE, dev_frames, n_frames, download=True)
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

    def reduce_min(x: float) -> float:
        return -reduce_max(-x)

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
        rows_per_step = reduce_sum(e2e["rows"]) / world / args.steps       # per rank, averaged over the ranks
        line["e2e"] = {"value": e_pairs / e_s, "unit": "frame-pairs/s",
                       "h2d_bytes_per_step": int(frame_bytes * sum(e2e_fps_all) / world),
                       "frames_per_step_per_rank": e2e_fps_all,
                       "sharding": ("equal shards" if len(set(e2e_fps_all)) == 1 else
                                    "shares of the step's frames proportional to each rank's measured e2e frame rate (two untimed calibration passes)"),
                       "d2h_bytes_per_step": int(rows_per_step * 16 + fps * max_corners * 8),
                       "wall_s": e2e["wall_s"], "clocks": e2e["clocks"],
                       "h2d_pinned_gbs_per_gpu": {"min": reduce_min(h2d_gbs), "max": reduce_max(h2d_gbs),
                                                  "sum": reduce_sum(h2d_gbs),
                                                  "note": "raw pinned host->device copy rate, every rank copying at once"},
                       "numa": numa}
        if "track" in e2e:
            line["e2e"]["track"] = e2e["track"]

    # ---- roofline of the dominant kernel (rank 0's numbers) ------------------------------
    peak, peak_src = read_peak_hbm()
    t = res["times"]
    fam = {k[:-3]: (t[k], t[k[:-3] + "_n"]) for k in t if k.endswith("_ms")}
    # dominant = most SM-time: the PnP solve is one 16-CTA cluster (16 of the SMs) on its own stream,
    # every other family fills the chip while it runs
    sm_share = {"pnp": 16.0 / 148.0}
    dominant = max(fam, key=lambda k: fam[k][0] * sm_share.get(k, 1.0))
    # With several detector / LK streams the event spans of the families overlap and include the wait for SM slots, so
    # the family that needs the most SM-time is taken from the committed launch list of this configuration
    # (profiles/roofline_traffic.json: isolated duration x share of the SMs the kernel holds) when there is one.
    try:
        prof_dom = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("%s/%s" % (args.config, args.motion), {})
        cand = {k: v.get("sm_us", 0.0) for k, v in prof_dom.items() if k in fam and fam[k][1]}
        if cand:
            dominant = max(cand, key=cand.get)
    except Exception:
        pass
    step_ms = sum(v[0] for v in fam.values())
    n_out = res["rows"] / max(res["pairs"], 1)
    per_launch = {
        "lk": 8 * lk_algorithmic_bytes(w, h, max_corners, int(n_out)),       # one launch = the 8 pairs of a frame
        "lk_tmpl": lk_algorithmic_bytes(w, h, max_corners, 0) // 2,          # the source frame's windows, once
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
    traffic = issue = None                                # per launch, from the committed ncu --set full capture
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        ent = prof.get("%s/%s" % (args.config, args.motion), {}).get(dominant)
        if ent:
            traffic = ent.get("bytes_per_launch")
            if ent.get("warp_inst_per_launch") and avg_ms > 0:
                # issue roof: one warp instruction per scheduler per clock, 4 schedulers per SM
                sm_mhz = res["clocks"].get("sm_mhz") or res["clocks"].get("sm_max_mhz") or 1965.0
                peak_inst = 148 * 4 * sm_mhz * 1e6
                ach_inst = ent["warp_inst_per_launch"] / (avg_ms * 1e-3)
                issue = {"bound": "issue", "warp_inst_per_launch": ent["warp_inst_per_launch"],
                         "achieved_ginst_s": ach_inst / 1e9, "peak_ginst_s": peak_inst / 1e9,
                         "frac": ach_inst / peak_inst,
                         # the live span covers the launch while it shares the chip with the other LK stream and the
                         # detector streams; alone (committed launch list) the kernel issues at this fraction of the roof
                         "frac_isolated": (ent["warp_inst_per_launch"] / (ent["isolated_us"] * 1e-6) / peak_inst
                                           if ent.get("isolated_us") else None),
                         "isolated_us": ent.get("isolated_us"),
                         "active_lanes_per_inst": ent.get("thread_inst_per_inst"), "source": ent.get("source")}
    except Exception:
        traffic = issue = None
    line["roofline"] = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                        "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": per_launch.get(dominant, 0),
                        "share_of_step": dom_ms / step_ms if step_ms else None,
                        "issue": issue,
                        "per_kernel": {k: {"ms_total": v[0], "launches": int(v[1]),
                                           "avg_ms": v[0] / max(v[1], 1),
                                           "achieved_gbs": (per_launch.get(k, 0) / (v[0] / max(v[1], 1) * 1e-3) / 1e9)
                                           if v[0] > 0 else 0.0,
                                           "frac_of_hbm_peak": (per_launch.get(k, 0) / (v[0] / max(v[1], 1) * 1e-3) / 1e9 / peak)
                                           if v[0] > 0 else 0.0}
                                       for k, v in fam.items() if v[1]}}

    # ---- the path's one collective (N > 1): stitch the poses this sweep produced ------------
    if world > 1:
        line["collective"] = run_collective(torch, dist, shard, res.get("poses", {}), start, args.steps * fps, world, rank)

    # ---- Refine at configs[4] scale and the plugin-surface number (N = 1) -------------------
    if world == 1 and not args.no_ba:
        try:
            line["ba"] = ba_block(ctx, capi, synth, args, w, h, max_corners, K, Rs, ts, verts, tris, dev_frames,
                                  frame_bytes, stride, n_frames, peak)
        except Exception as e:                            # the headline line must not be lost to a side block
            line["ba"] = {"error": repr(e)}
    if world == 1 and not args.no_plugin:
        try:
            line["plugin_e2e"] = plugin_block(args)
        except Exception as e:
            line["plugin_e2e"] = {"error": repr(e)}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) -----------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:                                              # the CPU arm gets every host core back
            os.sched_setaffinity(0, orig_affinity)
        except Exception:
            pass
        # the CPU arm in its own process (allocator state, CPU affinity and CUDA context do not leak into it)
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", args.config, "--motion",
               args.motion, "--steps", "1", "--warmup", "1", "--ref-frames-per-step", str(args.cpu_baseline_frames)]
        if not track:
            cmd.append("--no-track")
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(env, **CPU_ARM_ENV))
            ref = json.loads(out.stdout.strip().splitlines()[-1])
            line["cpu_baseline"] = ref["cpu_baseline"]
        except Exception as e:
            line["cpu_baseline"] = {"error": repr(e)}
    if world > 1:
        torch.cuda.synchronize()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))
    ctx.device_free(dev_frames)
    ctx.close()


def run_collective(torch, dist, shard, poses: dict, start: int, n_local: int, world: int, rank: int) -> dict:
    """All-gather of the packed pc_camera_state records (64 B per frame) of the frames each rank just tracked
    -- the stitch before the global refine -- on NCCL, timed with CUDA events on its
    own, and verified: every rank's slice of the result must be bit-equal to what that rank produced."""
    local = torch.zeros((n_local, shard.CAMERA_STATE_FLOATS), dtype=torch.float32)
    filled = 0
    for k in range(n_local):
        p = poses.get(start + k)
        if p is not None:
            local[k] = torch.from_numpy(p)
            filled += 1
    local = local.cuda()
    counts = [n_local] * world
    full = shard.allgather_trajectory(local, counts)            # warm-up (communicator set-up)
    torch.cuda.synchronize()
    dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    reps = 10
    ev[0].record()
    for _ in range(reps):
        full = shard.allgather_trajectory(local, counts)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # verification: exchange a checksum of every rank's local segment and compare with the gathered slices
    sums = torch.zeros(world, device="cuda", dtype=torch.float64)
    sums[rank] = local.double().sum()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ok = bool(torch.equal(full[rank * n_local:(rank + 1) * n_local], local))
    for r in range(world):
        ok = ok and bool(full[r * n_local:(r + 1) * n_local].double().sum() == sums[r])
    okt = torch.tensor([1.0 if ok else 0.0], device="cuda", dtype=torch.float64)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    return {"op": "all_gather of tracked poses (shard.allgather_trajectory, NCCL)", "frames_per_rank": n_local,
            "poses_filled_rank0": filled, "bytes_per_rank": n_local * 64, "ms": float(t.item()),
            "verified_against_per_rank_results": bool(okt.item() == 1.0), "in_timed_region": False}


def ba_block(ctx, capi, synth, args, w, h, max_corners, K, Rs, ts, verts, tris, dev_frames, frame_bytes, stride,
             n_frames, peak):
    """Refine at BASELINE configs[4] scale on this GPU: `--ba-frames` keyframes x max_corners tracks.  Analyze
    produces the flows (the reference's data path: Analyze writes them, Refine reads them), the trajectory is
    the ground truth + N(0, 0.2 deg / 0.5 % depth) on the interior frames (seed 1), loss Cauchy(1.0)."""
    F = min(args.ba_frames, n_frames)
    gftt = capi.default_gftt(max_corners=max_corners)
    kps, flows = {}, {}

    def take(r):
        kps[r["frame_id"]] = np.array(r["keypoints"], np.float32).reshape(-1, 2).copy()
        for (a, b, rows, idx, tgt, err) in r["pairs"]:
            flows[(a, b)] = (np.array(idx, np.uint32).copy(), np.array(tgt, np.float32).reshape(-1, 2).copy())

    ctx.analyze_begin(w, h, 0, F, gftt, capi.default_flow())
    for i in range(F):
        ctx.analyze_push(i, dev_frames + i * frame_bytes, stride, capi.PC_MEM_DEVICE)
        if ctx.analyze_pending() >= 4:
            take(ctx.analyze_pop(download=True, copy=True))
    while ctx.analyze_pending():
        take(ctx.analyze_pop(download=True, copy=True))
    ctx.analyze_end()
    ctx.mesh_set(verts, tris)
    edges = [(a, b, flows[(a, b)][0], flows[(a, b)][1]) for (a, b) in sorted(flows) if len(flows[(a, b)][0])]
    rows = int(sum(len(e[2]) for e in edges))
    t0 = time.perf_counter()
    ctx.ba_load([kps[k] for k in range(F)], edges, np.eye(4, dtype=np.float32), False, False)
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
            R = R @ (np.eye(3) + (np.sin(th) / th) * kx + ((1 - np.cos(th)) / th ** 2) * (kx @ kx))
            t = t + rng.normal(0, 0.005 * 4.0, 3)
        traj.append(capi.camera_state(K, R, t))

    def pose_err(tr):
        return max(float(np.abs(np.array(tr[i].t[:]) - np.array(truth[i].t[:])).max()) for i in range(F))

    bo = capi.default_bundle(loss_type=2, max_iterations=20)
    truth_cost = ctx.ba_cost(truth, bo)
    ctx.ba_cost(traj, bo)                                     # warm-up (fills the primitive-id cache)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.ba_cost(traj, bo)
    cost_s = (time.perf_counter() - t0) / reps
    ctx.ba_normal_equations(traj, bo)
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.ba_normal_equations(traj, bo)
    build_s = (time.perf_counter() - t0) / reps
    ctx.timing_read(reset=True)
    ctx.timing_enable(True)
    e0 = pose_err(traj)
    t0 = time.perf_counter()
    out, st = ctx.ba_solve(traj, bo)
    solve_s = time.perf_counter() - t0
    times = ctx.timing_read(reset=True)
    ctx.timing_enable(False)
    iters = max(int(st.iterations), 1)
    build_gbs = 24.0 * rows / build_s / 1e9                   # 24 B per residual row
    return {"workload": f"BASELINE configs[4]: {F} keyframes x {max_corners} tracks at {w}x{h}, {len(edges)} edges, "
                        f"{rows} residual rows, Cauchy(1.0), intrinsics fixed",
            "cost_eval_ms": 1e3 * cost_s, "normal_equation_build_ms": 1e3 * build_s,
            "note": "cost / build are whole C-ABI calls (trajectory upload, kernels, result download, one sync)",
            "build_rows_per_s": rows / build_s, "cost_rows_per_s": rows / cost_s,
            "lm_iterations": int(st.iterations), "invalid_steps": int(st.invalid_steps), "solve_wall_ms": 1e3 * solve_s,
            "lm_iteration_ms": 1e3 * solve_s / iters, "solve_gpu_kernel_ms": times.get("ba_ms"),
            "initial_cost": float(st.initial_cost), "final_cost": float(st.cost), "cost_at_ground_truth": truth_cost,
            "max_abs_t_err_before": e0, "max_abs_t_err_after": pose_err(out), "ba_load_wall_s": load_s,
            "roofline": {"kernel": "ba_build_kernel", "bound": "hbm", "algorithmic_bytes": 24 * rows,
                         "achieved": build_gbs, "peak": peak, "unit": "GB/s", "frac": build_gbs / peak}}


def plugin_block(args):
    """The Analyze pass through the reference's own entry point, in a fresh process like a user's (Blender's):
    `python bench.py --plugin-only`.  (Inside this process, next to a context that holds tens of GB, creating and
    destroying the analyze context alone costs over a second of cudaMalloc / cudaFree.)"""
    cmd = [sys.executable, os.path.abspath(__file__), "--plugin-only", "--config", args.config, "--motion", args.motion,
           "--plugin-frames", str(args.plugin_frames)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    if out.returncode != 0:
        return {"error": out.stderr[-500:]}
    return json.loads(out.stdout.strip().splitlines()[-1])


def plugin_only(args):
    """polychase_core.OpticalFlowThread, frames handed over as host numpy arrays one request at a time
    (opticalflow_thread.h:50-132), SQLite database written (database.cc:108-214).  Wall clock from the
    constructor to the terminal message."""
    import ctypes
    import shutil
    import tempfile
    from polychase_b200 import capi, synth
    from polychase_b200 import polychase_core as core
    w, h, max_corners, _ = CONFIGS[args.config]
    speed = synth.survey_speed(w) if args.motion == "survey" else 1.0
    F = args.plugin_frames
    ring = 32
    ctx = capi.Context(device=0, max_width=w, max_height=h, max_features=1024)        # renders the clip, then goes away
    ctx.synth_set_texture(synth.make_texture(w, h, seed=0))
    K = synth.intrinsics(w, h)
    scale = synth.plane_scale(w, 4.0)
    Rs, ts = synth.camera_path(ring, 4.0, 0, speed)
    dev = ctx.device_alloc(w * h * 3)
    host = []
    for i in range(ring):
        ctx.synth_render(synth.homography(K, Rs[i], ts[i], w, h, scale), dev, w * 3)
        ctx.synchronize()
        a = np.empty((h, w, 3), np.uint8)
        ctx.lib.pc_memcpy_d2h(ctx.h, a.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(dev), a.nbytes)
        host.append(a)
    ctx.device_free(dev)
    ctx.close()

    def image_of(idx):
        if idx < ring:
            return idx
        period = 2 * (ring - 1)
        m = idx % period
        return m if m < ring else period - m

    tmp = tempfile.mkdtemp()
    try:
        dbp = os.path.join(tmp, "clip.db")
        go = core.GFTTOptions()
        go.max_corners = max_corners
        errors = []
        t0 = time.perf_counter()
        th = core.OpticalFlowThread(core.VideoInfo(w, h, 1, F), dbp, go)
        t_first = t_last_req = None
        provide_s = 0.0
        while True:
            m = th.try_pop()
            if m is None:
                time.sleep(0.0002)
                continue
            if isinstance(m, bool):
                break
            if isinstance(m, core.OpticalFlowRequest):
                t1 = time.perf_counter()
                if t_first is None:
                    t_first = t1 - t0
                th.provide_frame(m.frame_id, host[image_of(m.frame_id - 1)])
                t_last_req = time.perf_counter()
                provide_s += t_last_req - t1
            elif isinstance(m, core.CppException):
                errors.append(m.what())
        th.join()
        dt = time.perf_counter() - t0
        pairs = 8 * F - 30
        print(json.dumps({
            "entry_point": "polychase_core.OpticalFlowThread (request / provide_frame hand-off) + SQLite database, "
                           "fresh process",
            "frames": F, "directed_pairs": pairs, "wall_s": dt, "value": pairs / dt, "unit": "frame-pairs/s",
            "frames_per_s": F / dt, "db_bytes": os.path.getsize(dbp), "errors": errors,
            "breakdown_s": {"until_first_request": t_first, "inside_provide_frame": provide_s,
                            "first_to_last_request": (t_last_req - t0 - t_first) if t_first is not None else None,
                            "after_last_request": dt - (t_last_req - t0) if t_last_req else None}}))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
