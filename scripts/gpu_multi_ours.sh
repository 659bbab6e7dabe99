#!/bin/bash
N=${1:-8}; TAG=${2:-rX}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -8; nproc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 \
    bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
cat $OUT/${TAG}_bench_n$N.json | cut -c1-1200; tail -3 $OUT/${TAG}_bench_n$N.err
