#!/bin/bash
# round 2, 8-GPU call: topology, NUMA placement probe of the pinned upload ring, default bench at N = 8
OUT=gpurun_out; mkdir -p $OUT
N=${1:-8}
nvidia-smi topo -m > $OUT/r2n_topo.txt 2>&1
for n in /sys/devices/system/node/node*; do echo "$n cpus $(cat $n/cpulist) mem $(grep MemTotal $n/meminfo)"; done >> $OUT/r2n_topo.txt 2>&1
grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status >> $OUT/r2n_topo.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    scripts/h2d_numa_probe.py > $OUT/r2n_h2d_numa_probe.json 2> $OUT/r2n_probe.err
cat $OUT/r2n_h2d_numa_probe.json | cut -c1-3000; tail -3 $OUT/r2n_probe.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 \
    bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r2n_bench_n$N.json 2> $OUT/r2n_bench_n$N.err
cat $OUT/r2n_bench_n$N.json | cut -c1-1500; tail -3 $OUT/r2n_bench_n$N.err
