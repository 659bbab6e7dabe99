"""Generates tests/golden/*.npz from OpenCV itself (cv2, pinned configuration) and the
reference's known-answer example.  Run here (the container with cv2 4.13.0); the fixtures are
committed so the GPU box never needs /root/reference or a particular cv2 build."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cvref, synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(OUT, exist_ok=True)
cvref.pin(1)
info = cvref.build_info_digest()

for name, (w, h, n) in {"a": (208, 144, 180), "b": (163, 121, 120)}.items():
    clip = synth.Clip(w, h, 5, seed=31 if name == "a" else 32)
    rng = np.random.default_rng(5)
    rgb0 = clip.rgb(0)
    rgb0[..., 1] = np.clip(rgb0[..., 1].astype(int) + rng.integers(-9, 9, rgb0.shape[:2]), 0, 255).astype(np.uint8)
    rgb1 = clip.rgb(4)
    g0, g1 = cvref.rgb2gray(rgb0), cvref.rgb2gray(rgb1)
    lv, dv = cvref.pyramid(g0)
    eig = cvref.min_eig(g0)
    kps, _ = cvref.gftt(g0, max_corners=n)
    kps_all, _ = cvref.gftt(g0, max_corners=0)
    nxt, st, err = cvref.lk(g0, g1, kps)
    np.savez_compressed(
        os.path.join(OUT, f"analyze_{name}.npz"), rgb0=rgb0, rgb1=rgb1, gray0=g0, gray1=g1,
        **{f"level{i}": l for i, l in enumerate(lv)}, **{f"deriv{i}": d for i, d in enumerate(dv)},
        eig=eig, kps=kps, kps_all=kps_all, lk_next=nxt, lk_status=st, lk_err=err, max_corners=n,
        cv2_info=np.array(info))
    print(name, w, h, "levels", len(lv), "kps", len(kps), len(kps_all), "tracked", int(st.sum()))
