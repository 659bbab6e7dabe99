#!/bin/bash
# N-GPU bench through torchrun, both arms (the driver's launch line).
N=${1:-2}; TAG=${2:-rX}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
cat $OUT/${TAG}_bench_n$N.json | cut -c1-900; tail -3 $OUT/${TAG}_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/${TAG}_ref_n$N.json 2> $OUT/${TAG}_ref_n$N.err
cat $OUT/${TAG}_ref_n$N.json | cut -c1-600; tail -3 $OUT/${TAG}_ref_n$N.err
