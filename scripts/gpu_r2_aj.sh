#!/bin/bash
# round 2, GPU call AJ: PnP cluster with 512 threads per CTA: parity, bench at 4K / 1080p / 720p, launch list
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_track_refine.py tests/test_gpu_ba_midsize.py tests/test_gpu_shard.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -4
for c in 4k 1080p 720p; do
  timeout 600 python bench.py --config $c --no-ba --no-plugin --no-cpu-baseline > $OUT/r2aj_bench_$c.json 2>> $OUT/r2aj_bench.err
  python -c "
import json; d=json.loads(open('$OUT/r2aj_bench_$c.json').read().strip().splitlines()[-1]); print('$c value', round(d['value']), 'e2e', round(d['e2e']['value']), 'pnp span', round(d['roofline']['per_kernel']['pnp']['avg_ms'],4), d['track'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/r2aj_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin > $OUT/r2aj_ncu_b.log 2>&1
python scripts/launch_summary.py $OUT/r2aj_launches.csv | head -6
