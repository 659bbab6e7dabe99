"""Builds the pybind11 module `polychase_core` (the reference's Python surface,
/root/reference/cpp/polychase_pybind.cc) in-tree with g++, linked against
lib/libpolychase_b200.so and the system libsqlite3.so.0.

    python -m polychase_b200.build_pybind [--force]
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
HOST = os.path.join(HERE, "csrc", "host")
OBJ_DIR = os.path.join(HERE, "build", "host")
EXT = sysconfig.get_config_var("EXT_SUFFIX")
OUT = os.path.join(HERE, "polychase_core" + EXT)


def build(force: bool = False) -> str:
    import pybind11
    from . import build as lib_build
    lib = lib_build.build()
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(HOST, "*.cc")))
    hdrs = glob.glob(os.path.join(HOST, "*.h")) + [os.path.join(HERE, "..", "include", "polychase_b200.h")]
    inc = ["-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"]]
    flags = ["-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wextra", "-Wno-unused-parameter", "-pthread"]
    objs = []
    for s in srcs:
        o = os.path.join(OBJ_DIR, os.path.basename(s) + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or any(os.path.getmtime(o) < os.path.getmtime(d) for d in [s] + hdrs):
            subprocess.check_call(["g++"] + flags + inc + ["-c", s, "-o", o])
    if force or not os.path.exists(OUT) or any(os.path.getmtime(OUT) < os.path.getmtime(o) for o in objs + [lib]):
        subprocess.check_call(["g++", "-shared", "-o", OUT] + objs +
                              ["-L" + os.path.dirname(lib), "-lpolychase_b200", "-l:libsqlite3.so.0", "-pthread",
                               "-Wl,-rpath,$ORIGIN/lib"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
