"""Builds libpolychase_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m polychase_b200.build [--force]

The shared object lands in polychase_b200/lib/ (git-ignored, shipped to the GPU box by
gpurun).  cudart is linked statically so the library has no CUDA runtime dependency
beyond the driver.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(LIB_DIR, "libpolychase_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]
# files whose float arithmetic must match the CPU reference bit for bit: no implicit FMA
NO_FMAD = {"lk.cu", "lk10.cu", "lk10q.cu", "mineig.cu"}


def sources():
    # csrc/host is the C++ host layer above the C ABI: it builds into the pybind11 module instead
    return sorted(glob.glob(os.path.join(CSRC, "kernels", "*.cu")) + glob.glob(os.path.join(CSRC, "abi", "*.cu")))


def headers():
    out = []
    for pat in ("kernels/*.h", "kernels/*.cuh", "abi/*.h"):
        out += glob.glob(os.path.join(CSRC, pat))
    out.append(os.path.join(HERE, "..", "include", "polychase_b200.h"))
    return out


def _compile(src: str, force: bool, log) -> str:
    obj = os.path.join(OBJ_DIR, os.path.basename(src) + ".o")
    deps = [src] + headers()
    if not force and os.path.exists(obj) and all(os.path.getmtime(obj) >= os.path.getmtime(d) for d in deps):
        return obj
    cmd = [NVCC] + ARCH + COMMON + ["-c", src, "-o", obj]
    if os.path.basename(src) in NO_FMAD:
        cmd.append("-fmad=false")
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append((src, r.stdout + r.stderr))
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sources()
    log = []
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, log), srcs))
    if verbose:
        for src, text in log:
            print(f"== {os.path.basename(src)}\n{text}")
    if (force or not os.path.exists(LIB)
            or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)):
        cmd = [NVCC] + ARCH + ["-shared", "-cudart", "static", "-o", LIB] + objs + ["-ldl", "-lpthread"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv))
