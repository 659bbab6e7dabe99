#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=$1
show() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    pk=d['roofline']['per_kernel']
    print(sys.argv[1], round(d['value']), round(d.get('e2e',{}).get('value',0)), round(d['ms_per_step'],2), {k:round(v['ms_total']/d['steps'],2) for k,v in pk.items()})
except Exception as e: print(sys.argv[1], 'ERR', e)
P
}
for depth in ${DEPTHS:-4 6 8}; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --depth $depth > $OUT/${TAG}_depth$depth.json 2>> $OUT/${TAG}.err; show $OUT/${TAG}_depth$depth.json
done
[ -n "$SKIP_SMALL" ] || timeout 300 python bench.py --steps 10 --warmup 3 --config 1080p > $OUT/${TAG}_1080p.json 2>> $OUT/${TAG}.err; show $OUT/${TAG}_1080p.json
[ -n "$SKIP_SMALL" ] || timeout 300 python bench.py --steps 10 --warmup 3 --config 720p --no-cpu-baseline > $OUT/${TAG}_720p.json 2>> $OUT/${TAG}.err; show $OUT/${TAG}_720p.json
tail -5 $OUT/${TAG}.err
