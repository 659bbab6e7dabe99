#!/bin/bash
# round 2, GPU call L: second LK stream A/B, full GPU suite, launch list
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py 2>&1 | tail -25 > $OUT/r2l_tests.log
tail -6 $OUT/r2l_tests.log
for v in "PC_LK_STREAMS=2" "PC_LK_STREAMS=1" "PC_LK_STREAMS=2 PC_LK_QUEUE=1" "PC_LK_STREAMS=2 PC_H2D_SPLIT=2"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 600 python bench.py --no-ba --no-plugin --no-cpu-baseline > $OUT/r2l_bench_${tag}.json 2>> $OUT/r2l_bench.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2l_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "e2e", round(d.get("e2e", {}).get("value", 0)), "ms/step", round(d["ms_per_step"], 2))
        print("  per_kernel", {k: round(v["avg_ms"], 4) for k, v in d["roofline"]["per_kernel"].items()})
    except Exception as e:
        print(f, "ERR", repr(e))
PY
tail -c 600 $OUT/r2l_bench.err
