#!/bin/bash
# round 2, 8-GPU call AK: upload slices per frame at N = 8 (32 vs 16 vs 8 concurrent copy streams on the box)
OUT=gpurun_out; mkdir -p $OUT
N=8
for sp in 2 1; do
  PC_H2D_SPLIT=$sp timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$sp \
      bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r2ak_bench_n${N}_split$sp.json 2> $OUT/r2ak_bench_n${N}_split$sp.err
  python -c "
import json; d=json.loads(open('$OUT/r2ak_bench_n${N}_split$sp.json').read().strip().splitlines()[-1]); print('N=$N split=$sp value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('frames_per_step_per_rank'))"
done
