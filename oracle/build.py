"""Builds the oracle's C restatement (gcc) -- TEST INFRASTRUCTURE.

Output: oracle/_build/liboracle_restate.so (git-ignored via *.so; it does travel to the
GPU box with the gpurun snapshot).  The reference itself cannot be compiled here
(needs Eigen3, oneTBB, Embree 4, OpenCV C++ headers, spdlog, sqlite3.h, cmake+vcpkg;
SURVEY.md section 0.3), so there is no oracle/_ref.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle_restate.so")
SRC = [os.path.join(HERE, "restate.c"), os.path.join(HERE, "track_port.c")]


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(LIB)
            and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in SRC)):
        return LIB
    cmd = ["gcc", "-O2", "-std=gnu11", "-ffp-contract=off", "-fno-fast-math", "-mfma",
           "-shared", "-fPIC", "-o", LIB] + SRC + ["-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
