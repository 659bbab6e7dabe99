"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    name = row["Kernel Name"].split("(")[0][-60:]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':62s} {'n':>5s} {'total_us':>10s} {'avg_us':>9s} share")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:62s} {v[0]:5d} {v[1]:10.1f} {v[1] / v[0]:9.1f} {v[1] / tot:.3f}")
