"""LK iteration statistics of the benchmark clip (CPU, oracle trace): how unbalanced are the six keypoints that share a
warp of the lock-step kernel, and what would re-ordering / deferral schemes recover?  Analysis tooling (it drives the
oracle's C restatement with its per-level iteration trace); the numbers quoted in DESIGN.md section 4 come from here.

    python scripts/lk_iteration_stats.py > profiles/r2_lk_iteration_stats.txt      (about two minutes, no GPU)
"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import gftt as ogftt  # noqa: E402
from oracle import restate, synth  # noqa: E402
from polychase_b200 import synth as psynth  # noqa: E402

W, H, N = 3840, 2160, 8000


def ratio(t):
    """warp rounds of the lock-step kernel (sum over levels of the max over a warp's six keypoints) / ideal"""
    n = len(t) // 6 * 6
    t = t[:n]
    return t.reshape(-1, 6, 4).max(1).sum() / (t.sum() / 6)


def deferral(t, J0, T):
    """lock-step main kernel in which, after J0 iterations of a level, the last T still-iterating pentads of a warp leave
    for a perfectly balanced second stage: (main rounds, second-stage rounds, deferred items), rounds / ideal"""
    n = len(t) // 6 * 6
    t = t[:n]
    main = rem = nd = 0
    for g in range(0, n, 6):
        c = t[g:g + 6]
        alive = np.ones(6, bool)
        for L in (3, 2, 1, 0):
            cnt = np.where(alive, c[:, L], 0)
            j = 0
            while True:
                it = cnt > j
                k = it.sum()
                if k == 0:
                    break
                if j >= J0 and k <= T:
                    for p in np.nonzero(it)[0]:
                        rem += cnt[p] - j + c[p][:L].sum()
                        alive[p] = False
                        nd += 1
                    break
                j += 1
            main += j
    ideal = t.sum() / 6
    return main / ideal, rem / 6 / ideal, nd


def main():
    clip = synth.Clip(W, H, 12, seed=0, speed=psynth.survey_speed(W))
    g = {k: clip.gray(k) for k in (0, 1, 4, 8)}
    kps = ogftt.gftt_from_eig(restate.min_eig(g[0], mode=3), max_corners=N)
    pyr = {k: restate.pyramid(g[k], 3) for k in g}
    lib = restate.lib()
    res = {}
    for d in (1, 4, 8):
        tr = np.zeros((len(kps), 4), np.int32)
        lib.orc_lk_set_iter_trace(tr.ctypes.data_as(C.c_void_p))
        restate.lk(pyr[0], pyr[d], kps)
        lib.orc_lk_set_iter_trace(None)
        res[d] = tr.copy()
    print(f"4K, {N} keypoints of frame 0, survey-conformant clip; iterations per level from the oracle (levels 0..3)")
    for d, t in res.items():
        n = len(t) // 6 * 6
        mean = t[:n].mean(0)
        grp = t[:n].reshape(-1, 6, 4).max(1).mean(0)
        tot = t.sum(1)
        print(f"pair 0 -> {d}: mean iterations per level {np.round(mean, 2)}, mean of the max over six {np.round(grp, 2)}, "
              f"total mean {mean.sum():.1f}, lock-step / ideal {ratio(t):.3f}; keypoints with >= 40 iterations: {(tot >= 40).sum()}")
        cx, cy = (kps[:, 0] // 64).astype(np.int64), (kps[:, 1] // 64).astype(np.int64)
        print(f"   ordered by 64-px cell: {ratio(t[np.argsort(cy * 1000 + cx, kind='stable')]):.3f};  "
              f"ordered by the (unknowable) total: {ratio(t[np.argsort(tot, kind='stable')]):.3f}")
        for (J0, T) in ((4, 1), (6, 1), (4, 2)):
            m, s, nd = deferral(t, J0, T)
            print(f"   deferral J0={J0} T={T}: main {m:.3f} + balanced second stage {s:.3f} = {m + s:.3f}  ({nd} items deferred)")
    t4, t8 = res[4], res[8]
    print(f"correlation of a keypoint's total between 0->4 and 0->8: {np.corrcoef(t4.sum(1), t8.sum(1))[0, 1]:.2f}; "
          f"0->8 ordered by the 0->4 totals: {ratio(t8[np.argsort(t4.sum(1), kind='stable')]):.3f}")


if __name__ == "__main__":
    main()
