"""Oracle restatement vs the committed golden vectors (tests/golden/*.npz, produced by cv2
4.13.0 with scripts/make_golden.py).  CPU only; does not need cv2."""
import glob
import os

import numpy as np
import pytest

from oracle import gftt as ogftt
from oracle import restate

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "analyze_*.npz")))


def _u32(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("path", GOLD)
def test_restatement_reproduces_golden(path):
    z = np.load(path)
    assert np.array_equal(restate.rgb2gray(z["rgb0"]), z["gray0"])
    assert np.array_equal(restate.rgb2gray(z["rgb1"]), z["gray1"])
    lv = restate.pyramid(z["gray0"], 3)
    for i in range(4):
        assert np.array_equal(lv[i], z[f"level{i}"])
        assert np.array_equal(restate.scharr(lv[i]), z[f"deriv{i}"].reshape(lv[i].shape + (2,)))
    assert np.array_equal(_u32(restate.min_eig(z["gray0"], 1)), _u32(z["eig"]))
    assert np.array_equal(ogftt.detect(z["gray0"], 1, max_corners=int(z["max_corners"])), z["kps"])
    assert np.array_equal(ogftt.detect(z["gray0"], 1, max_corners=0), z["kps_all"])
    nx, st, er = restate.lk(lv, restate.pyramid(z["gray1"], 3), z["kps"])
    assert np.array_equal(st, z["lk_status"])
    assert np.array_equal(_u32(nx), _u32(z["lk_next"]))
    ok = st == 1
    assert np.array_equal(_u32(er[ok]), _u32(z["lk_err"][ok]))


def test_goldens_present():
    assert len(GOLD) >= 2
