"""Writes profiles/roofline_traffic.json: per kernel family, per launch -- DRAM traffic and warp instructions from an
`ncu --set full` capture (default cache control: L2 flushed before every kernel, i.e. cold-cache traffic) and the
isolated duration from the `--metrics gpu__time_duration.sum` launch list of the same command.

    python scripts/make_roofline_traffic.py gpurun_out/r2h_prof.ncu-rep gpurun_out/r2m_launches.csv 4k/survey \
        profiles/r2_h_ncu_all_kernels.txt profiles/r2_m_launch_summary.txt

bench.py looks the entry of its dominant family up under "<config>/<motion>".
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

rep, launches, key, src_ncu, src_list = sys.argv[1:6]
FAMILIES = [("lk_tmpl", r"lk10q?_template_kernel|spatial_order_kernel"), ("lk", r"lk10q?_kernel|lk10q_err_kernel|lk_kernel"), ("compact", r"lk_compact_kernel"),
            ("gray_pyr", r"gray_l1_tma|l2_l3_tma|pad_border|rgb_to_gray|pyr_down"), ("min_eig", r"min_eig_kernel|init_cell_max"),
            ("select", r"nms_candidates|greedy_suppress|compact_top|select_rank"), ("raycast", r"raycast_"), ("pnp", r"pnp_lm_kernel")]


def family(name):
    for f, pat in FAMILIES:
        if re.search(pat, name):
            return f
    return None


raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {n: hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
                                 "smsp__thread_inst_executed_per_inst_executed.ratio")}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per_kernel = {}                                    # first captured launch of every kernel
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].split("::")[-1].replace("void ", "")
    if name in per_kernel:
        continue
    b = sum(float(r[col[c]]) * scale.get(units[col[c]], 1) for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    per_kernel[name] = {"bytes": b, "inst": float(r[col["smsp__inst_executed.sum"]]), "thr": float(r[col["smsp__thread_inst_executed_per_inst_executed.ratio"]]) * float(r[col["smsp__inst_executed.sum"]])}
iso = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader([l for l in open(launches) if not l.startswith("==")]):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    name = row["Kernel Name"].split("(")[0].split("::")[-1].replace("void ", "")
    iso[name][0] += 1
    iso[name][1] += v
out = {}
for fam, _ in FAMILIES:
    ks = [k for k in per_kernel if family(k) == fam]
    if not ks:
        continue
    inst = sum(per_kernel[k]["inst"] for k in ks)
    out[fam] = {"kernels": sorted(ks), "bytes_per_launch": int(sum(per_kernel[k]["bytes"] for k in ks)),
                "warp_inst_per_launch": int(inst),
                "thread_inst_per_inst": round(sum(per_kernel[k]["thr"] for k in ks) / max(inst, 1), 2),
                "isolated_us": round(sum(iso[k][1] / iso[k][0] for k in ks if iso[k][0]), 1),
                # SM-time: kernels that are one 16-CTA cluster hold 16 of the 148 SMs while they run
                "sm_us": round(sum(iso[k][1] / iso[k][0] * (16.0 / 148.0 if re.search(r"greedy_suppress|pnp_lm", k) else 1.0)
                                   for k in ks if iso[k][0]), 1),
                "source": f"{src_ncu} (cold cache), {src_list}"}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "roofline_traffic.json")
try:
    doc = json.load(open(path))
    doc = {k: v for k, v in doc.items() if "/" in k or k == "_comment"}
except Exception:
    doc = {}
doc["_comment"] = ("Per kernel family and launch: DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) and warp instructions from an "
                   "ncu --set full capture of `python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-ba --no-plugin` (L2 flushed before "
                   "every kernel: cold-cache traffic; write-backs still in L2 at kernel end are not counted), isolated_us from the launch list "
                   "of the same command.  bench.py picks its dominant family by sm_us (isolated duration x share of the SMs the kernel holds) and copies that entry into roofline.traffic / roofline.issue.")
doc[key] = out
json.dump(doc, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
