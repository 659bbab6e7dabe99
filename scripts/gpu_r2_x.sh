#!/bin/bash
# round 2, GPU call X: spatial keypoint order for the lock-step LK kernel: parity, A/B bench lines, launch list, ncu of LK
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_analyze.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py tests/test_gpu_shard.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -15 > $OUT/r2x_tests.log
tail -5 $OUT/r2x_tests.log
for v in "PC_LK_SPATIAL=1" "PC_LK_SPATIAL=0"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 600 python bench.py --no-ba --no-plugin --no-cpu-baseline > $OUT/r2x_bench_${tag}.json 2>> $OUT/r2x_bench.err
  python -c "
import json; d=json.loads(open('$OUT/r2x_bench_${tag}.json').read().strip().splitlines()[-1]); print('$v value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v['avg_ms'],4) for k,v in d['roofline']['per_kernel'].items()})"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/r2x_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin > $OUT/r2x_ncu_b.log 2>&1
python scripts/launch_summary.py $OUT/r2x_launches.csv > $OUT/r2x_launch_summary.txt; cat $OUT/r2x_launch_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:lk10_kernel|lk10_template|spatial_order" -s 30 -c 3 \
    -o $OUT/r2x_prof -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin > $OUT/r2x_ncu_full.log 2>&1
