#!/bin/bash
# ncu only: launch list of one bench step + full capture of the named kernels.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_prof.sh r5 "lk10_kernel|nms_candidates" 40 8'
TAG=${1:-rX}; KERNELS=${2:-"lk10_kernel"}; SKIP=${3:-40}; COUNT=${4:-6}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/${TAG}_ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$KERNELS" -s $SKIP -c $COUNT \
    -o $OUT/${TAG}_prof -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -5
