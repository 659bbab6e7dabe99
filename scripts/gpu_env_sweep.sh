#!/bin/bash
# Bench the same build under several environment settings (one line per setting).
#   gpurun --timeout 900 -- 'bash scripts/gpu_env_sweep.sh r11 "PC_LK_BLOCKS_PER_SM=3" "PC_LK_BLOCKS_PER_SM=2"'
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
i=0
for setting in "default=1" "$@"; do
  env $setting timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/${TAG}_sweep_$i.json 2>> $OUT/${TAG}_sweep.err
  python - "$setting" $OUT/${TAG}_sweep_$i.json <<'P'
import json,sys
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
pk=d['roofline']['per_kernel']
print(sys.argv[1], round(d['value']), round(d['ms_per_step'],2), {k:round(v['ms_total']/d['steps'],2) for k,v in pk.items()})
P
  i=$((i+1))
done 2>&1 | tee $OUT/${TAG}_sweep.txt
