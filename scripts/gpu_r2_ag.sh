#!/bin/bash
# round 2, GPU call AG: final build of the round (NMS per-tile append, sub-bin ranking, spatial LK order, 3 detector streams): full GPU suite, smoke, default bench line, reference arm, launch list
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $OUT/r2ag_tests.log
tail -8 $OUT/r2ag_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > $OUT/r2ag_bench_4k.json 2> $OUT/r2ag_bench.err
tail -c 300 $OUT/r2ag_bench.err
python -c "
import json; d=json.loads(open('$OUT/r2ag_bench_4k.json').read().strip().splitlines()[-1]); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'plugin', d['plugin_e2e'].get('value'), 'ba ms', d['ba']['solve_wall_ms'], 'cpu', d['cpu_baseline']['value']); print(d['roofline']['issue'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r2ag_ref_4k.json 2>> $OUT/r2ag_bench.err
cut -c1-200 $OUT/r2ag_ref_4k.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/r2ag_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin > $OUT/r2ag_ncu_b.log 2>&1
python scripts/launch_summary.py $OUT/r2ag_launches.csv > $OUT/r2ag_launch_summary.txt; cat $OUT/r2ag_launch_summary.txt
for c in 1080p 720p; do
  timeout 600 python bench.py --config $c --no-ba --no-plugin > $OUT/r2ag_bench_$c.json 2>> $OUT/r2ag_bench.err
  python -c "
import json; d=json.loads(open('$OUT/r2ag_bench_$c.json').read().strip().splitlines()[-1]); print('$c value', round(d['value']), 'e2e', round(d['e2e']['value']), 'cpu', d['cpu_baseline']['value'])"
done
