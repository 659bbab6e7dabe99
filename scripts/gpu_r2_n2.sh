#!/bin/bash
# round 2, 2-GPU call: the NCCL tests of the C ABI (all-gather, edge-sharded refine) and the bench at N = 2
OUT=gpurun_out; mkdir -p $OUT
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_shard.py -m gpu -q 2>&1 | tail -15 > $OUT/r2p_tests_n$N.log; tail -5 $OUT/r2p_tests_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 \
    bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r2p_bench_n$N.json 2> $OUT/r2p_bench_n$N.err
python -c "
import json; d=json.loads(open('$OUT/r2p_bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('frames_per_step_per_rank'), d['e2e']['h2d_pinned_gbs_per_gpu'], d.get('collective'))"
tail -3 $OUT/r2p_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29528 \
    bench.py --gpus $N --steps 10 --warmup 3 --equal-shards > $OUT/r2p_bench_n${N}_equal.json 2>> $OUT/r2p_bench_n$N.err
python -c "
import json; d=json.loads(open('$OUT/r2p_bench_n${N}_equal.json').read().strip().splitlines()[-1]); print('N=$N equal shards: value', round(d['value']), 'e2e', round(d['e2e']['value']))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29529 \
    bench.py --impl reference --gpus $N --steps 1 --warmup 0 > $OUT/r2p_ref_n$N.json 2>> $OUT/r2p_bench_n$N.err
cut -c1-300 $OUT/r2p_ref_n$N.json
