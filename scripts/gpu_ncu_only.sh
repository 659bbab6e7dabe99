#!/bin/bash
# ncu --set full capture of the named kernels only.
#   gpurun --timeout 900 -- 'bash scripts/gpu_ncu_only.sh r14 "pad_border|lk_compact" 30 10'
TAG=${1:-rX}; KERNELS=${2:-"lk10_kernel"}; SKIP=${3:-40}; COUNT=${4:-6}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k "regex:$KERNELS" -s $SKIP -c $COUNT \
    -o $OUT/${TAG}_prof -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -3
