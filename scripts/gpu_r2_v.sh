#!/bin/bash
# round 2, GPU call V: full GPU suite (incl. the 4K track + refine property test), smoke, default bench line
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $OUT/r2v_tests.log
tail -8 $OUT/r2v_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > $OUT/r2v_bench_4k.json 2> $OUT/r2v_bench.err
tail -c 300 $OUT/r2v_bench.err
python -c "
import json; d=json.loads(open('$OUT/r2v_bench_4k.json').read().strip().splitlines()[-1]); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'plugin', d['plugin_e2e'].get('value'), 'ba ms', d['ba']['solve_wall_ms'], 'cpu', d['cpu_baseline']['value']); print(d['roofline']['issue'])"
