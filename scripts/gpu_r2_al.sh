#!/bin/bash
# round 2, 2-GPU call AL: final check of the N > 1 bench path (both arms) and the 2-GPU tests with the final build
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/r2al_bench_n2.json 2> $OUT/r2al_bench_n2.err
python -c "
import json; d=json.loads(open('$OUT/r2al_bench_n2.json').read().strip().splitlines()[-1]); print('N=2 value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('frames_per_step_per_rank'), d['collective']['verified_against_per_rank_results'])"
tail -2 $OUT/r2al_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 \
    bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>> $OUT/r2al_bench_n2.err | cut -c1-160
