"""Pins the oracle restatement (oracle/restate.c, oracle/gftt.py) against OpenCV itself (cv2),
live, on fresh inputs -- the third-party arithmetic the reference calls.
CPU only."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import cvref, restate, synth
from oracle import gftt as ogftt


@pytest.fixture(autouse=True)
def _pin():
    cvref.pin(1)


def _u32(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("w,h", [(643, 487), (320, 240), (101, 77), (64, 48)])
def test_gray_pyramid_scharr_bit_exact(w, h):
    rng = np.random.default_rng(w + h)
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    g = cvref.rgb2gray(rgb)
    assert np.array_equal(g, restate.rgb2gray(rgb))
    lv, dv = cvref.pyramid(g)
    assert len(lv) == ogftt.num_pyramid_levels(w, h, 10, 3)
    mine = restate.pyramid(g, 3)
    for i in range(len(lv)):
        assert np.array_equal(lv[i], mine[i])
        assert np.array_equal(dv[i].reshape(lv[i].shape + (2,)), restate.scharr(mine[i]))


@pytest.mark.parametrize("w,h", [(640, 480), (643, 487), (100, 75), (105, 60), (37, 29), (208, 144), (112, 60), (250, 40)])
def test_min_eig_bit_exact(w, h):
    tex = synth.make_texture(w, h, seed=3)
    assert np.array_equal(_u32(cvref.min_eig(tex)), _u32(restate.min_eig(tex, 1)))
    # the independent-box variant (what the CUDA kernel computes) differs in a few ppm at most
    d = _u32(cvref.min_eig(tex)) != _u32(restate.min_eig(tex, 3))
    assert d.mean() < 5e-5


def test_min_eig_plain_path():
    tex = synth.make_texture(333, 222, seed=9)
    cv2.setUseOptimized(False)
    try:
        e = cvref.min_eig(tex)
    finally:
        cv2.setUseOptimized(True)
    assert np.array_equal(_u32(e), _u32(restate.min_eig(tex, 0)))


@pytest.mark.parametrize("w,h,mc", [(640, 480, 0), (643, 487, 700), (320, 240, 100)])
def test_detector_equals_cv2_based(w, h, mc):
    g = synth.Clip(w, h, 1, seed=11).gray(0)
    a, _ = cvref.gftt(g, max_corners=mc)
    b = ogftt.detect(g, 1, max_corners=mc)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("min_distance", [1.0, 3.0, 5.4, 7.5, 0.5])
def test_detector_min_distance_variants(min_distance):
    g = synth.Clip(200, 150, 1, seed=12).gray(0)
    a, _ = cvref.gftt(g, min_distance=min_distance, max_corners=0)
    b = ogftt.detect(g, 1, min_distance=min_distance, max_corners=0)
    assert np.array_equal(a, b)


def test_detector_flat_and_empty():
    flat = np.full((64, 80), 77, np.uint8)
    assert len(ogftt.detect(flat, 1)) == 0
    a, _ = cvref.gftt(flat)
    assert len(a) == 0


@pytest.mark.parametrize("w,h,n,skip", [(640, 480, 800, 4), (333, 251, 400, 1), (160, 120, 150, 8)])
def test_lk_bit_exact(w, h, n, skip):
    clip = synth.Clip(w, h, skip + 1, seed=5)
    g1, g2 = clip.gray(0), clip.gray(skip)
    pts, _ = cvref.gftt(g1, max_corners=n)
    rng = np.random.default_rng(0)
    extra = np.stack([rng.uniform(-3, w + 3, 100), rng.uniform(-3, h + 3, 100)], 1).astype(np.float32)
    pts = np.concatenate([pts, extra]).astype(np.float32)
    nx, st, er = cvref.lk(g1, g2, pts)
    nlev = ogftt.num_pyramid_levels(w, h, 10, 3)
    nx2, st2, er2 = restate.lk(restate.pyramid(g1, 3)[:nlev], restate.pyramid(g2, 3)[:nlev], pts)
    assert np.array_equal(st, st2)
    assert np.array_equal(_u32(nx), _u32(nx2))
    ok = st == 1
    assert np.array_equal(_u32(er[ok]), _u32(er2[ok]))


@pytest.mark.parametrize("win,lvl", [(7, 2), (15, 3), (5, 1)])
def test_lk_other_windows(win, lvl):
    w, h = 320, 240
    clip = synth.Clip(w, h, 3, seed=9)
    g1, g2 = clip.gray(0), clip.gray(2)
    pts, _ = cvref.gftt(g1, max_corners=200)
    nx, st, er = cvref.lk(g1, g2, pts, win=win, max_level=lvl)
    nlev = ogftt.num_pyramid_levels(w, h, win, lvl)
    nx2, st2, er2 = restate.lk(restate.pyramid(g1, lvl)[:nlev], restate.pyramid(g2, lvl)[:nlev], pts, win=win)
    assert np.array_equal(st, st2)
    assert np.array_equal(_u32(nx), _u32(nx2))
