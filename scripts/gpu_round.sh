#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list, one ncu --set full capture.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh r3 "lk10_kernel|pnp_lm_kernel"'
TAG=${1:-rX}
KERNELS=${2:-"lk10_kernel|pnp_lm_kernel|min_eig_kernel"}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L; nproc
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/${TAG}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --diag > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
PC_LK_STREAM=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_1stream.json 2>> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench_1stream.json
if [ -n "$KERNELS" ] && [ "$KERNELS" != "none" ]; then
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/${TAG}_ncu_b.log 2>&1
# one full capture of the named kernels (skip the warm-up launches)
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$KERNELS" -s ${NCU_SKIP:-40} -c ${NCU_COUNT:-8} \
    -o $OUT/${TAG}_prof -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
fi
ls -la $OUT | tail -8
