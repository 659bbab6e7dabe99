#!/bin/bash
# round 2, third 8-GPU call: e2e with calibrated ingest-proportional shards vs equal shards
OUT=gpurun_out; mkdir -p $OUT
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 \
    bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r2u_bench_n$N.json 2> $OUT/r2u_bench_n$N.err
python -c "
import json; d=json.loads(open('$OUT/r2u_bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('frames_per_step_per_rank'), d['e2e']['h2d_pinned_gbs_per_gpu'])"
tail -3 $OUT/r2u_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29528 \
    bench.py --gpus $N --steps 10 --warmup 3 --equal-shards > $OUT/r2u_bench_n${N}_equal.json 2>> $OUT/r2u_bench_n$N.err
python -c "
import json; d=json.loads(open('$OUT/r2u_bench_n${N}_equal.json').read().strip().splitlines()[-1]); print('N=$N equal shards: value', round(d['value']), 'e2e', round(d['e2e']['value']))"
