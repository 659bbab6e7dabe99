#!/bin/bash
# round 2, GPU call Y: 2 / 3 / 4 detector streams
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_analyze.py tests/test_gpu_shard.py -m gpu -q -x 2>&1 | tail -3
for v in "PC_DET_STREAMS=2" "PC_DET_STREAMS=3" "PC_DET_STREAMS=4" "PC_DET_STREAMS=3 PC_FRAMES_DEPTH=24"; do
  tag=$(echo "$v" | tr ' =' '__')
  extra=""; case "$v" in *DEPTH*) extra="--depth 24";; esac
  env $v timeout 600 python bench.py --no-ba --no-plugin --no-cpu-baseline $extra > $OUT/r2y_bench_${tag}.json 2>> $OUT/r2y_bench.err
  python -c "
import json; d=json.loads(open('$OUT/r2y_bench_${tag}.json').read().strip().splitlines()[-1]); print('$v value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v['avg_ms'],4) for k,v in d['roofline']['per_kernel'].items()})"
done
tail -c 300 $OUT/r2y_bench.err
