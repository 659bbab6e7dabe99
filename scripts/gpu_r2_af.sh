#!/bin/bash
# round 2, GPU call AF: ranking kernel with per-bin sub-bin sort (min_eig unsplit): parity, launch list, bench
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_analyze.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -m gpu -q -x 2>&1 | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/r2af_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin > $OUT/r2af_ncu_b.log 2>&1
python scripts/launch_summary.py $OUT/r2af_launches.csv > $OUT/r2af_launch_summary.txt; cat $OUT/r2af_launch_summary.txt
timeout 600 python bench.py --no-ba --no-plugin --no-cpu-baseline > $OUT/r2af_bench_4k.json 2> $OUT/r2af_bench.err
python -c "
import json; d=json.loads(open('$OUT/r2af_bench_4k.json').read().strip().splitlines()[-1]); print('value', round(d['value']), 'e2e', round(d['e2e']['value']))"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:select_rank_emit|compact_top" -s 10 -c 2 \
    -o $OUT/r2af_prof -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin > $OUT/r2af_ncu_full.log 2>&1
