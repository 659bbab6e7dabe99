#!/bin/bash
# round 2, GPU call B: full GPU suite, the restructured bench on both clips, launch list of one step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -25 > gpurun_out/r2b_tests.log
cat gpurun_out/r2b_tests.log | tail -12
timeout 900 python bench.py > gpurun_out/r2b_bench_4k_survey.json 2> gpurun_out/r2b_bench.err
tail -c 600 gpurun_out/r2b_bench.err
timeout 600 python bench.py --motion r1 --no-ba --no-plugin > gpurun_out/r2b_bench_4k_r1.json 2>> gpurun_out/r2b_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2b_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin > gpurun_out/r2b_ncu_b.log 2>&1
python scripts/launch_summary.py gpurun_out/r2b_launches.csv > gpurun_out/r2b_launch_summary.txt 2>&1
cat gpurun_out/r2b_launch_summary.txt | head -40
python - <<'PY'
import json
for f in ("r2b_bench_4k_survey", "r2b_bench_4k_r1"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 2), d.get("track"), d.get("cpu_baseline", {}).get("value"))
        print("  per_kernel", {k: round(v["avg_ms"], 4) for k, v in d["roofline"]["per_kernel"].items()})
        print("  ba", d.get("ba")); print("  plugin", d.get("plugin_e2e"))
    except Exception as e:
        print(f, "ERR", e)
PY
