#!/bin/bash
# round 2, 8-GPU call: the final build at N = 8, 4, 2 (the driver's launch line), both legs
OUT=gpurun_out; mkdir -p $OUT
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N \
      bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r2ai_bench_n$N.json 2> $OUT/r2ai_bench_n$N.err
  python -c "
import json; d=json.loads(open('$OUT/r2ai_bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('frames_per_step_per_rank'), 'h2d sum', round(d['e2e']['h2d_pinned_gbs_per_gpu']['sum'],1), d.get('collective',{}).get('verified_against_per_rank_results'))"
  tail -2 $OUT/r2ai_bench_n$N.err
done
