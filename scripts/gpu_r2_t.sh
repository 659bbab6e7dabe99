#!/bin/bash
# round 2, 2-GPU call T: calibrated ingest-proportional shards (code path check at N = 2), plugin block with the copy pool
OUT=gpurun_out; mkdir -p $OUT
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 \
    bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r2t_bench_n$N.json 2> $OUT/r2t_bench_n$N.err
python -c "
import json; d=json.loads(open('$OUT/r2t_bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('frames_per_step_per_rank'), d['e2e'].get('sharding'))"
tail -3 $OUT/r2t_bench_n$N.err
for i in 1 2; do timeout 300 python bench.py --plugin-only > $OUT/r2t_plugin_$i.json 2>> $OUT/r2t_bench_n$N.err; cat $OUT/r2t_plugin_$i.json; done
timeout 300 python bench.py --plugin-only --plugin-frames 1024 > $OUT/r2t_plugin_1024.json 2>> $OUT/r2t_bench_n$N.err; cat $OUT/r2t_plugin_1024.json
timeout 600 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_analyze.py -m gpu -q -x 2>&1 | tail -4
