#!/bin/bash
# round 2, GPU call C: on-device refine (tests + bench_ba), plugin breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_track_refine.py tests/test_gpu_ba_midsize.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2c_tests.log
tail -15 gpurun_out/r2c_tests.log
timeout 300 python scripts/bench_ba.py --frames 200 --max-iterations 20 > gpurun_out/r2c_ba_200.json 2> gpurun_out/r2c_ba.err
timeout 300 python scripts/bench_ba.py --frames 1000 --max-iterations 20 > gpurun_out/r2c_ba_1000.json 2>> gpurun_out/r2c_ba.err
tail -3 gpurun_out/r2c_ba.err
cat gpurun_out/r2c_ba_200.json gpurun_out/r2c_ba_1000.json
timeout 600 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-ba > gpurun_out/r2c_bench_plugin.json 2> gpurun_out/r2c_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2c_bench_plugin.json').read().strip().splitlines()[-1]); print(d.get('plugin_e2e'))"
timeout 300 python scripts/bench_dropin.py --frames 128 --no-refine > gpurun_out/r2c_dropin.jsonl 2>> gpurun_out/r2c_bench.err; cat gpurun_out/r2c_dropin.jsonl
