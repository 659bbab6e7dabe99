#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/probe.py <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from polychase_b200 import capi
ctx = capi.Context(max_width=1920, max_height=1088, max_features=2048)
rng = np.random.default_rng(1)
rgb = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
try:
    ctx.upload_rgb(1, rgb)
    print("upload ok", ctx.read_level(1, 1)[:2, :8])
except Exception as e:
    print("ERR", e)
PY
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/probe.py > gpurun_out/r2e_sanitizer.log 2>&1
head -60 gpurun_out/r2e_sanitizer.log
