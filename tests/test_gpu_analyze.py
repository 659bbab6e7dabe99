"""GPU parity tests of the analyze path (K1-K9) through the C ABI, against the oracle
restatement (oracle/restate.c, itself pinned bit-exact to cv2 in test_oracle_vs_cv2.py)."""
import numpy as np
import pytest

from oracle import restate, synth
from oracle import gftt as ogftt

pytestmark = pytest.mark.gpu


def _u32(a):
    return np.ascontiguousarray(a).view(np.uint32)


# (640, 480) .. (1916, 1082): the fused TMA path (pyramid_tma.cu: width a multiple of 4), with edge tiles of every
# kind -- widths / heights that are not multiples of the 128 x 32 block, odd level sizes, a single tile row;
# (643, 487), (100, 75): the generic kernels (gray_pyr.cu)
@pytest.mark.parametrize("w,h", [(640, 480), (643, 487), (1280, 720), (100, 75), (644, 483), (1284, 722), (1916, 1082),
                                 (300, 37), (136, 35), (292, 90)])
def test_gray_pyramid_bit_exact(ctx_small, w, h):
    rng = np.random.default_rng(w * 7 + h)
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ctx_small.upload_rgb(1, rgb)
    gray = restate.rgb2gray(rgb)
    levels = restate.pyramid(gray, 3)
    n = ctx_small.num_levels(1)
    assert n == ogftt.num_pyramid_levels(w, h, 10, 3)
    for L in range(n):
        got = ctx_small.read_level(1, L)
        assert got.shape == levels[L].shape
        assert np.array_equal(got, levels[L]), f"level {L} differs"


@pytest.mark.parametrize("w,h", [(640, 480), (643, 487), (1280, 720), (101, 77)])
def test_min_eig_map(ctx_small, w, h):
    tex = synth.make_texture(w, h, seed=3)
    ctx_small.upload_gray(2, tex)
    got = ctx_small.min_eig_map(2, w, h)
    want_indep = restate.min_eig(tex, mode=3)     # same op order, 3x3 sums formed independently
    assert np.array_equal(_u32(got), _u32(want_indep))
    want_cv = restate.min_eig(tex, mode=1)        # OpenCV's running column sum (bit-exact vs cv2)
    diff = _u32(got) != _u32(want_cv)
    assert diff.mean() < 2e-5, f"{diff.sum()} pixels differ from the cv2-exact restatement"


@pytest.mark.parametrize("w,h,max_corners", [(640, 480, 0), (1280, 720, 2000), (643, 487, 500), (1920, 1080, 4000)])
def test_detector_matches_oracle(ctx_small, w, h, max_corners):
    clip = synth.Clip(w, h, 2, seed=11)
    g = clip.gray(0)
    from polychase_b200 import capi
    ctx_small.upload_gray(3, g)
    got = ctx_small.detect(3, capi.default_gftt(max_corners=max_corners))
    eig = restate.min_eig(g, mode=3)
    want = ogftt.gftt_from_eig(eig, max_corners=max_corners)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    # and against the cv2-exact eig map: same list unless one of the ~2 ppm pixels matters
    want_cv = ogftt.gftt_from_eig(restate.min_eig(g, mode=1), max_corners=max_corners)
    assert np.array_equal(got, want_cv)


@pytest.mark.parametrize("w,h,n,skip", [(640, 480, 1500, 4), (1280, 720, 2000, 8), (333, 251, 600, 1),
                                          (1920, 1080, 4000, 8)])          # the last: BASELINE configs[1] shape
def test_lk_bit_exact(ctx_small, w, h, n, skip):
    clip = synth.Clip(w, h, skip + 1, seed=5)
    g1, g2 = clip.gray(0), clip.gray(skip)
    eig = restate.min_eig(g1, mode=1)
    pts = ogftt.gftt_from_eig(eig, max_corners=n)
    rng = np.random.default_rng(0)
    extra = np.stack([rng.uniform(-3, w + 3, 200), rng.uniform(-3, h + 3, 200)], 1).astype(np.float32)
    pts = np.concatenate([pts, extra]).astype(np.float32)
    ctx_small.upload_gray(10, g1)
    ctx_small.upload_gray(11, g2)
    ctx_small.set_keypoints(10, pts)
    nxt, st, err = ctx_small.lk_raw(10, 11)
    L1, L2 = restate.pyramid(g1, 3), restate.pyramid(g2, 3)
    nlev = ogftt.num_pyramid_levels(w, h, 10, 3)
    wn, ws, we = restate.lk(L1[:nlev], L2[:nlev], pts)
    assert np.array_equal(st, ws)
    assert np.array_equal(_u32(nxt), _u32(wn))
    ok = ws == 1
    assert np.array_equal(_u32(err[ok]), _u32(we[ok]))
    # the filtered rows the reference stores (opticalflow.cc:139-147)
    idx, tgt, e = ctx_small.lk_pair(10, 11)
    assert np.array_equal(idx, np.nonzero(ok)[0].astype(np.uint32))
    assert np.array_equal(_u32(tgt), _u32(wn[ok]))
    assert np.array_equal(_u32(e), _u32(we[ok]))


def test_lk_other_window_sizes(ctx_small):
    from polychase_b200 import capi
    w, h = 320, 240
    clip = synth.Clip(w, h, 3, seed=9)
    g1, g2 = clip.gray(0), clip.gray(2)
    pts = ogftt.gftt_from_eig(restate.min_eig(g1, 1), max_corners=300)
    ctx_small.upload_gray(20, g1)
    ctx_small.upload_gray(21, g2)
    ctx_small.set_keypoints(20, pts)
    for win, lvl in [(7, 2), (15, 3), (16, 1), (5, 3)]:
        fo = capi.default_flow(window_size=win, max_level=lvl)
        ctx_small.upload_gray(20, g1, fo)
        ctx_small.upload_gray(21, g2, fo)
        ctx_small.set_keypoints(20, pts)
        nxt, st, err = ctx_small.lk_raw(20, 21, fo)
        nlev = ogftt.num_pyramid_levels(w, h, win, lvl)
        L1, L2 = restate.pyramid(g1, lvl), restate.pyramid(g2, lvl)
        wn, ws, we = restate.lk(L1[:nlev], L2[:nlev], pts, win=win)
        assert np.array_equal(st, ws), f"win {win}"
        assert np.array_equal(_u32(nxt), _u32(wn)), f"win {win}"


def test_streaming_analyzer_matches_pairwise(ctx_small):
    """pc_analyze_* must produce exactly the rows GenerateOpticalFlowDatabase writes:
    keypoints per frame and the 8F-30 directed pairs (opticalflow.cc:237-316)."""
    from polychase_b200 import capi
    w, h, F = 320, 240, 12
    clip = synth.Clip(w, h, F, seed=21, first_frame=1)
    go = capi.default_gftt(max_corners=300)
    frames = {k: clip.rgb(k) for k in range(1, F + 1)}
    ctx_small.analyze_begin(w, h, 1, F, go)
    got_kps, got_pairs = {}, {}
    for k in range(1, F + 1):
        ctx_small.analyze_push(k, frames[k])
        if ctx_small.analyze_pending() >= 3:
            r = ctx_small.analyze_pop()
            got_kps[r["frame_id"]] = r["keypoints"]
            for (a, b, rows, idx, tgt, err) in r["pairs"]:
                got_pairs[(a, b)] = (idx, tgt, err)
    while ctx_small.analyze_pending():
        r = ctx_small.analyze_pop()
        got_kps[r["frame_id"]] = r["keypoints"]
        for (a, b, rows, idx, tgt, err) in r["pairs"]:
            got_pairs[(a, b)] = (idx, tgt, err)
    ctx_small.analyze_end()
    assert len(got_kps) == F
    expect_pairs = [(a, a + d) for a in range(1, F + 1) for d in (-8, -4, -2, -1, 1, 2, 4, 8) if 1 <= a + d <= F]
    assert sorted(got_pairs) == sorted(expect_pairs)
    grays = {k: restate.rgb2gray(frames[k]) for k in frames}
    pyr = {k: restate.pyramid(grays[k], 3) for k in frames}
    for k in frames:
        want = ogftt.gftt_from_eig(restate.min_eig(grays[k], 3), max_corners=300)
        assert np.array_equal(got_kps[k], want)
    for (a, b) in expect_pairs:
        wn, ws, we = restate.lk(pyr[a], pyr[b], got_kps[a])
        ok = ws == 1
        idx, tgt, err = got_pairs[(a, b)]
        assert np.array_equal(idx, np.nonzero(ok)[0].astype(np.uint32))
        assert np.array_equal(_u32(tgt), _u32(wn[ok]))
        assert np.array_equal(_u32(err), _u32(we[ok]))


@pytest.mark.parametrize("kw", [
    dict(min_distance=1.0, max_corners=0), dict(min_distance=3.0, max_corners=700),
    dict(min_distance=5.4, max_corners=0), dict(min_distance=7.5, max_corners=400),      # R > 4: generic disc scan
    dict(min_distance=0.5, max_corners=0), dict(min_distance=0.5, max_corners=900),      # < 1: nothing is suppressed
    dict(min_distance=30.0, max_corners=500),     # the strongest 4*max_corners keep too few: second pass over everything
    dict(quality_level=0.4, max_corners=8000),    # fewer corners than max_corners
    dict(quality_level=0.001, max_corners=1000),
    dict(grid_rows=1, grid_cols=1, max_corners=800), dict(grid_rows=2, grid_cols=3, max_corners=0),
    dict(grid_rows=8, grid_cols=8, max_corners=1200),
])
def test_detector_option_variants(ctx_small, kw):
    """Every branch of the selection stage (gftt.cc:38-192) against the restated detector."""
    from polychase_b200 import capi
    w, h = 803, 601
    g = synth.Clip(w, h, 1, seed=17).gray(0)
    ctx_small.upload_gray(30, g)
    got = ctx_small.detect(30, capi.default_gftt(**kw))
    want = ogftt.detect(g, 3, **kw)
    assert got.shape == want.shape
    assert np.array_equal(got, want)


@pytest.mark.parametrize("w,h", [(16, 16), (17, 23), (64, 48)])
def test_tiny_and_flat_frames(ctx_small, w, h):
    """Smallest accepted frames, and a frame without a single corner: zero keypoints, zero flow rows."""
    from polychase_b200 import capi
    rng = np.random.default_rng(w + h)
    g = rng.integers(0, 256, (h, w), dtype=np.uint8)
    ctx_small.upload_gray(40, g)
    got = ctx_small.detect(40, capi.default_gftt(max_corners=50))
    want = ogftt.detect(g, 3, max_corners=50)
    assert np.array_equal(got, want)
    flat = np.full((h, w), 127, np.uint8)
    ctx_small.upload_gray(41, flat)
    assert ctx_small.detect(41, capi.default_gftt(max_corners=50)).shape == (0, 2)
    ctx_small.upload_gray(42, g)
    idx, tgt, err = ctx_small.lk_pair(41, 42)
    assert len(idx) == 0 and len(tgt) == 0 and len(err) == 0


def test_streaming_with_a_featureless_frame_and_short_clips(ctx_small):
    """Ragged input: a flat frame in the middle of a clip (no keypoints: its outgoing pairs have zero rows,
    its incoming pairs are tracked into a constant image), and clips shorter than the +-8 window."""
    from polychase_b200 import capi
    w, h, F = 320, 240, 5
    clip = synth.Clip(w, h, F, seed=33, first_frame=0)
    frames = {k: clip.rgb(k) for k in range(F)}
    frames[2] = np.full((h, w, 3), 90, np.uint8)
    go = capi.default_gftt(max_corners=250)
    ctx_small.analyze_begin(w, h, 0, F, go)
    kps, pairs = {}, {}
    for k in range(F):
        ctx_small.analyze_push(k, frames[k])
    while ctx_small.analyze_pending():
        r = ctx_small.analyze_pop()
        kps[r["frame_id"]] = np.array(r["keypoints"]).copy()
        for (a, b, rows, idx, tgt, err) in r["pairs"]:
            pairs[(a, b)] = (np.array(idx).copy(), np.array(tgt).copy(), np.array(err).copy())
    ctx_small.analyze_end()
    expect = sorted((a, a + d) for a in range(F) for d in (-8, -4, -2, -1, 1, 2, 4, 8) if 0 <= a + d < F)
    assert sorted(pairs) == expect
    assert kps[2].shape == (0, 2)
    grays = {k: restate.rgb2gray(frames[k]) for k in frames}
    pyr = {k: restate.pyramid(grays[k], 3) for k in frames}
    for (a, b) in expect:
        idx, tgt, err = pairs[(a, b)]
        if a == 2:
            assert len(idx) == 0
            continue
        wn, ws, we = restate.lk(pyr[a], pyr[b], kps[a])
        ok = ws == 1
        assert np.array_equal(idx, np.nonzero(ok)[0].astype(np.uint32)), (a, b)
        assert np.array_equal(_u32(tgt), _u32(wn[ok])), (a, b)
        assert np.array_equal(_u32(err), _u32(we[ok])), (a, b)
    # a one-frame "clip": keypoints, no pairs
    ctx_small.analyze_begin(w, h, 7, 1, go)
    ctx_small.analyze_push(7, frames[0])
    r = ctx_small.analyze_pop()
    ctx_small.analyze_end()
    assert r["frame_id"] == 7 and len(r["pairs"]) == 0 and len(r["keypoints"]) == len(kps[0])


def test_streaming_with_preset_keypoints_on_the_borders(ctx_small):
    """Keypoints handed in (ReadOrGenerateKeypoints, opticalflow.cc:168-178) that sit on, near and outside
    the image borders, at an odd frame size: the cached-template LK path must agree with the oracle bit
    for bit (masked derivative taps, windows hanging over the edge, points outside a level)."""
    from polychase_b200 import capi
    w, h, F = 333, 251, 10
    clip = synth.Clip(w, h, F, seed=41, first_frame=0)
    frames = {k: clip.rgb(k) for k in range(F)}
    rng = np.random.default_rng(7)
    presets = {}
    for k in range(F):
        inside = np.stack([rng.uniform(0, w - 1, 300), rng.uniform(0, h - 1, 300)], 1)
        edge = np.concatenate([
            np.stack([rng.uniform(-6, 6, 60), rng.uniform(-6, h + 6, 60)], 1),
            np.stack([rng.uniform(w - 7, w + 6, 60), rng.uniform(-6, h + 6, 60)], 1),
            np.stack([rng.uniform(-6, w + 6, 60), rng.uniform(-6, 6, 60)], 1),
            np.stack([rng.uniform(-6, w + 6, 60), rng.uniform(h - 7, h + 6, 60)], 1),
            np.array([[0.0, 0.0], [w - 1.0, h - 1.0], [-30.0, 40.0], [w + 25.0, h + 25.0], [4.5, 4.5], [w - 5.5, h - 5.5]]),
        ])
        presets[k] = np.concatenate([inside, edge]).astype(np.float32)
    ctx_small.analyze_begin(w, h, 0, F, capi.default_gftt(max_corners=100))
    for k in range(F):
        ctx_small.analyze_preset_keypoints(k, presets[k])
    kps, pairs = {}, {}

    def take(r):
        kps[r["frame_id"]] = np.array(r["keypoints"]).copy()
        for (a, b, rows, idx, tgt, err) in r["pairs"]:
            pairs[(a, b)] = (np.array(idx).copy(), np.array(tgt).copy(), np.array(err).copy())

    for k in range(F):
        ctx_small.analyze_push(k, frames[k])
        if ctx_small.analyze_pending() >= 4:
            take(ctx_small.analyze_pop())
    while ctx_small.analyze_pending():
        take(ctx_small.analyze_pop())
    ctx_small.analyze_end()
    grays = {k: restate.rgb2gray(frames[k]) for k in frames}
    nlev = ogftt.num_pyramid_levels(w, h, 10, 3)
    pyr = {k: restate.pyramid(grays[k], 3)[:nlev] for k in frames}
    assert len(pairs) == 8 * F - 30
    for k in range(F):
        assert np.array_equal(kps[k], presets[k])
    for (a, b), (idx, tgt, err) in pairs.items():
        wn, ws, we = restate.lk(pyr[a], pyr[b], presets[a])
        ok = ws == 1
        assert np.array_equal(idx, np.nonzero(ok)[0].astype(np.uint32)), (a, b)
        assert np.array_equal(_u32(tgt), _u32(wn[ok])), (a, b)
        assert np.array_equal(_u32(err), _u32(we[ok])), (a, b)


def _stream_rows(ctx, w, h, frames, max_corners):
    from polychase_b200 import capi
    F = len(frames)
    ctx.analyze_begin(w, h, 0, F, capi.default_gftt(max_corners=max_corners))
    kps, pairs = {}, {}

    def take(r):
        kps[r["frame_id"]] = np.array(r["keypoints"]).copy()
        for (a, b, rows, idx, tgt, err) in r["pairs"]:
            pairs[(a, b)] = (np.array(idx).copy(), np.array(tgt).copy(), np.array(err).copy())

    for k in range(F):
        ctx.analyze_push(k, frames[k])
        if ctx.analyze_pending() >= 4:
            take(ctx.analyze_pop())
    while ctx.analyze_pending():
        take(ctx.analyze_pop())
    ctx.analyze_end()
    return kps, pairs


@pytest.mark.parametrize("env", [{"PC_LK_QUEUE": "1"}, {"PC_LK_QUEUE": "1", "PC_LK_BUDGET": "24"},
                                 {"PC_LK_QUEUE": "1", "PC_LK_BUDGET": "100"},
                                 {"PC_DET_STREAMS": "1", "PC_LK_STREAMS": "1"}, {"PC_DET_STREAMS": "1"}, {"PC_LK_STREAMS": "1"},
                                 {"PC_LK_STREAM": "0"}])
def test_lk_work_queue_schedules_agree(ctx_small, env, monkeypatch):
    """The 10x10 LK can run as a work queue (lk10q.cu, PC_LK_QUEUE=1: a pentad takes the next (pair, keypoint) as
    soon as it is done) -- a change of schedule only.  Every schedule (the default lock-step kernel = the session
    context, queue with resident blocks, queue with blocks that leave after 24 or 100 items) must give the same
    rows, bit for bit; one pair is also compared with the oracle.  The same holds for the stream layout: the session
    context runs two detector and two LK streams (frames / batches alternate, csrc/abi/capi.cu); one of each, and
    everything on one stream (PC_LK_STREAM=0), must give identical rows."""
    from polychase_b200 import capi
    w, h, F = 640, 480, 11
    clip = synth.Clip(w, h, F, seed=77, first_frame=0)
    frames = [clip.rgb(k) for k in range(F)]
    kps0, pairs0 = _stream_rows(ctx_small, w, h, frames, 1500)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    other = capi.Context(max_width=w, max_height=h, max_features=2048)
    try:
        kps1, pairs1 = _stream_rows(other, w, h, frames, 1500)
    finally:
        other.close()
    assert sorted(pairs0) == sorted(pairs1) and len(pairs0) == 8 * F - 30
    for k in range(F):
        assert np.array_equal(kps0[k], kps1[k])
    for key in pairs0:
        for x, y in zip(pairs0[key], pairs1[key]):
            assert np.array_equal(_u32(x), _u32(y)), key
    a, b = 10, 2
    pyr_a = restate.pyramid(restate.rgb2gray(frames[a]), 3)
    pyr_b = restate.pyramid(restate.rgb2gray(frames[b]), 3)
    wn, ws, we = restate.lk(pyr_a, pyr_b, kps0[a])
    ok = ws == 1
    idx, tgt, err = pairs0[(a, b)]
    assert np.array_equal(idx, np.nonzero(ok)[0].astype(np.uint32))
    assert np.array_equal(_u32(tgt), _u32(wn[ok]))
    assert np.array_equal(_u32(err), _u32(we[ok]))
