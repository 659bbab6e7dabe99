#!/bin/bash
# round 2, GPU call A: the new parity tests + the refine drift question (old vs survey motion)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_env.txt 2>&1
nproc >> gpurun_out/r2a_env.txt; numactl --hardware >> gpurun_out/r2a_env.txt 2>&1; nvidia-smi topo -m >> gpurun_out/r2a_env.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/r2a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
tail -30 gpurun_out/r2a_tests.log
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py 2>&1 | tail -40 > gpurun_out/r2a_tests_all.log
timeout 300 python scripts/bench_ba.py --frames 200 --max-iterations 20 --motion r1 > gpurun_out/r2a_ba_r1_20.json 2> gpurun_out/r2a_ba.err
timeout 300 python scripts/bench_ba.py --frames 200 --max-iterations 100 --motion r1 > gpurun_out/r2a_ba_r1_100.json 2>> gpurun_out/r2a_ba.err
timeout 300 python scripts/bench_ba.py --frames 200 --max-iterations 20 --motion survey > gpurun_out/r2a_ba_survey_20.json 2>> gpurun_out/r2a_ba.err
timeout 300 python scripts/bench_ba.py --frames 200 --max-iterations 100 --motion survey > gpurun_out/r2a_ba_survey_100.json 2>> gpurun_out/r2a_ba.err
cat gpurun_out/r2a_ba_*.json
