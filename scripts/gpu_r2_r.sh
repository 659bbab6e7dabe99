#!/bin/bash
# round 2, GPU call R: gray + level-1 kernel (DP4A gray, packed 16-bit pyrDown passes, fixed thread mapping), 16-byte side bands in pad_border: parity, bench, launch list
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_analyze.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -m gpu -q -x 2>&1 | tail -15 > $OUT/r2r_tests.log
tail -5 $OUT/r2r_tests.log
timeout 600 python bench.py --no-ba --no-plugin --no-cpu-baseline > $OUT/r2r_bench_4k.json 2> $OUT/r2r_bench.err
python -c "
import json; d=json.loads(open('$OUT/r2r_bench_4k.json').read().strip().splitlines()[-1]); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v['avg_ms'],4) for k,v in d['roofline']['per_kernel'].items()}); print({k:d['roofline'][k] for k in d['roofline'] if k!='per_kernel'})"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/r2r_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin > $OUT/r2r_ncu_b.log 2>&1
python scripts/launch_summary.py $OUT/r2r_launches.csv > $OUT/r2r_launch_summary.txt; cat $OUT/r2r_launch_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:gray_l1_tma|l2_l3_tma|pad_border" -s 30 -c 3 --cache-control none \
    -o $OUT/r2r_prof -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin > $OUT/r2r_ncu_full.log 2>&1
