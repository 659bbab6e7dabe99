#!/bin/bash
# round 2, GPU call I: LK work-queue kernel (lk10q.cu): parity tests, A/B against the lock-step kernel, block budgets,
# launch list + ncu --set full of the queue kernel
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_analyze.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py tests/test_gpu_shard.py -m gpu -q -x 2>&1 | tail -25 > $OUT/r2i_tests.log
tail -8 $OUT/r2i_tests.log
for v in "PC_LK_QUEUE=0" "PC_LK_QUEUE=1" "PC_LK_BUDGET=48" "PC_LK_BUDGET=96" "PC_LK_BUDGET=192"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 600 python bench.py --no-ba --no-plugin --no-cpu-baseline > $OUT/r2i_bench_${tag}.json 2>> $OUT/r2i_bench.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2i_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "e2e", round(d.get("e2e", {}).get("value", 0)), "ms/step", round(d["ms_per_step"], 2))
        print("  per_kernel", {k: round(v["avg_ms"], 4) for k, v in d["roofline"]["per_kernel"].items()})
    except Exception as e:
        print(f, "ERR", repr(e))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/r2i_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin > $OUT/r2i_ncu_b.log 2>&1
python scripts/launch_summary.py $OUT/r2i_launches.csv > $OUT/r2i_launch_summary.txt; cat $OUT/r2i_launch_summary.txt
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k "regex:lk10q" -s 60 -c 6 \
    -o $OUT/r2i_prof -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin > $OUT/r2i_ncu_full.log 2>&1
ls -la $OUT | tail -5
