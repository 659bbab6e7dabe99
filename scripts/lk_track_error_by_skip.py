"""How good are the LK tracks of a synthetic 4K clip, by frame skip, for the two camera motions bench.py knows?

    python scripts/lk_track_error_by_skip.py > profiles/r2_lk_track_error_by_skip.txt      (on a GPU box)

For every directed pair (a -> b) of a 17-frame 4K clip the tracked target of each status-1 keypoint is compared with
its true position H_b H_a^-1 x (the scene is a plane, so the truth is a homography).  `--motion r1` is round 1's
camera path (up to 89 px of image motion per 8 frames), `survey` the path SURVEY.md section 8d specifies (<~ 20 px).
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from polychase_b200 import capi, synth  # noqa: E402

W, H, N, F = 3840, 2160, 8000, 17


def run(speed, label, ctx):
    clip = synth.Clip(W, H, F, seed=0, speed=speed)
    ctx.analyze_begin(W, H, 0, F, capi.default_gftt(max_corners=N))
    kps, pairs = {}, {}

    def take(r):
        kps[r["frame_id"]] = np.array(r["keypoints"]).copy()
        for (a, b, rows, idx, tgt, err) in r["pairs"]:
            pairs[(a, b)] = (np.array(idx).copy(), np.array(tgt).copy())

    for k in range(F):
        ctx.analyze_push(k, clip.rgb(k))
        if ctx.analyze_pending() >= 4:
            take(ctx.analyze_pop())
    while ctx.analyze_pending():
        take(ctx.analyze_pop())
    ctx.analyze_end()
    print(f"motion {label} (time scale {speed:.3f}): 4K, {N} features/frame, {F} frames, {len(pairs)} directed pairs")
    print(f"  {'skip':>4s} {'pairs':>5s} {'tracked':>8s} {'median px':>10s} {'p90 px':>8s} {'>1 px':>7s} {'>3 px':>7s} {'true motion px (median / max)':>30s}")
    for d in (1, 2, 4, 8):
        errs, mot, n_src = [], [], 0
        for (a, b), (idx, tgt) in pairs.items():
            if abs(a - b) != d:
                continue
            Hab = clip.homography(b) @ np.linalg.inv(clip.homography(a))
            src = kps[a][idx].astype(np.float64)
            p = np.concatenate([src, np.ones((len(src), 1))], 1) @ Hab.T
            truth = p[:, :2] / p[:, 2:3]
            errs.append(np.linalg.norm(truth - tgt, axis=1))
            mot.append(np.linalg.norm(truth - src, axis=1))
            n_src += len(kps[a])
        e, m = np.concatenate(errs), np.concatenate(mot)
        print(f"  {d:4d} {len(errs):5d} {len(e) / n_src:8.3f} {np.median(e):10.3f} {np.percentile(e, 90):8.3f} "
              f"{(e > 1).mean():7.3f} {(e > 3).mean():7.3f} {np.median(m):18.1f} / {m.max():.1f}")


if __name__ == "__main__":
    with capi.Context(max_width=W, max_height=H, max_features=N, pipeline_depth=4) as ctx:
        run(1.0, "r1", ctx)
        run(synth.survey_speed(W), "survey", ctx)
