#!/bin/bash
# Launch list with warm caches (ncu --cache-control none): per-kernel durations as the pipeline sees them.
TAG=${1:-rX}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $OUT/${TAG}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file $OUT/${TAG}_launches_warm.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/${TAG}_ncu_b.log 2>&1
python scripts/launch_summary.py $OUT/${TAG}_launches_warm.csv | tee $OUT/${TAG}_launch_summary_warm.txt
