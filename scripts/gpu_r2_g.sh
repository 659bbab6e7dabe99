#!/bin/bash
# round 2, GPU call G: full GPU suite (no -x), default bench, H2D experiments (slices on several copy engines,
# write-combined pinned memory), launch list
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1
numactl --hardware >> gpurun_out/r2g_topo.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py 2>&1 | tail -40 > gpurun_out/r2g_tests.log
tail -15 gpurun_out/r2g_tests.log
timeout 900 python bench.py > gpurun_out/r2g_bench_4k.json 2> gpurun_out/r2g_bench.err
tail -c 400 gpurun_out/r2g_bench.err
for v in "PC_H2D_SPLIT=2" "PC_H2D_SPLIT=4" "PC_PINNED_WC=1" "PC_PINNED_WC=1 PC_H2D_SPLIT=2"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 600 python bench.py --no-ba --no-plugin --no-cpu-baseline > gpurun_out/r2g_bench_${tag}.json 2>> gpurun_out/r2g_bench.err
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2g_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin > gpurun_out/r2g_ncu_b.log 2>&1
python - <<'PY'
import json, csv, collections, glob
for f in sorted(glob.glob("gpurun_out/r2g_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "e2e", d.get("e2e", {}).get("value"), "ms/step", round(d["ms_per_step"], 2))
        print("  per_kernel", {k: round(v["avg_ms"], 4) for k, v in d["roofline"]["per_kernel"].items()})
        if "ba" in d: print("  ba", d["ba"])
        if "plugin_e2e" in d: print("  plugin", d.get("plugin_e2e"))
        if "cpu_baseline" in d: print("  cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(f, "ERR", repr(e))
lines = [l for l in open("gpurun_out/r2g_launches.csv") if not l.startswith("==")]
agg = collections.defaultdict(lambda: collections.defaultdict(float)); cnt = collections.Counter()
for row in csv.DictReader(lines):
    try: v = float(row["Metric Value"].replace(",", ""))
    except Exception: continue
    name = row["Kernel Name"].split("(")[0][-40:]
    m = row["Metric Name"]; u = row["Metric Unit"]
    if m.startswith("gpu__time"): v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v; cnt[name] += 1
    else: v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    agg[name][m] += v
print("%-42s %5s %9s %9s %9s" % ("kernel", "n", "avg_us", "rd_MB", "wr_MB"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    n = cnt[k]
    print("%-42s %5d %9.1f %9.2f %9.2f" % (k, n, v["gpu__time_duration.sum"] / n, v["dram__bytes_read.sum"] / n / 1e6, v["dram__bytes_write.sum"] / n / 1e6))
PY
