"""CUDA path vs the committed cv2 golden vectors (tests/golden/*.npz) through the C ABI."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "analyze_*.npz")))


def _u32(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("path", GOLD)
def test_cuda_reproduces_cv2_golden(ctx_small, path):
    from polychase_b200 import capi
    z = np.load(path)
    ctx_small.upload_rgb(1, z["rgb0"])
    ctx_small.upload_rgb(2, z["rgb1"])
    for i in range(4):
        assert np.array_equal(ctx_small.read_level(1, i), z[f"level{i}"])
    kps = ctx_small.detect(1, capi.default_gftt(max_corners=int(z["max_corners"])))
    assert np.array_equal(kps, z["kps"])
    assert np.array_equal(ctx_small.detect(1, capi.default_gftt(max_corners=0)), z["kps_all"])
    ctx_small.set_keypoints(1, z["kps"])
    nxt, st, err = ctx_small.lk_raw(1, 2)
    assert np.array_equal(st, z["lk_status"])
    assert np.array_equal(_u32(nxt), _u32(z["lk_next"]))
    ok = st == 1
    assert np.array_equal(_u32(err[ok]), _u32(z["lk_err"][ok]))
