"""Summarise an .ncu-rep (raw metrics + hottest source lines) -- used to produce profiles/*.txt."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units = r[0], r[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_eligible.avg.per_cycle_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for row in r[2:]:
    name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("kernel:", name[:90])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w} = {row[i]} {units[i]}")
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(row[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    print("  stalls per issue: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))
    pipes = []
    for i, h in enumerate(hdr):
        if h.startswith("sm__inst_executed_pipe_") and h.endswith(".avg.pct_of_peak_sustained_active"):
            try:
                pipes.append((float(row[i]), h[len("sm__inst_executed_pipe_"):-len(".avg.pct_of_peak_sustained_active")]))
            except ValueError:
                pass
    print("  pipes % of peak: " + ", ".join(f"{n} {v:.1f}" for v, n in sorted(pipes, reverse=True)[:6]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = None
agg = []
for x in rows:
    if len(x) >= 2 and x[0] == "File Path":
        cur = x[1].split("/")[-1]
        continue
    if len(x) < 8 or x[0] in ("Line No", "Function Name"):
        continue
    if x[0] != "" and x[2] == "-":
        try:
            agg.append((int(x[7]), int(x[4]), cur, int(x[0]), x[1].strip()[:90]))
        except ValueError:
            pass
tot = sum(a[0] for a in agg) or 1
print(f"hottest source lines (of {tot} warp instructions):")
for a in sorted(agg, reverse=True)[:top]:
    print(f"  {a[0] / tot * 100:5.1f}% inst  {a[1]:6d} samples  {a[2]}:{a[3]:<4d} {a[4]}")
tot_s = sum(a[1] for a in agg) or 1
print(f"source lines with the most warp samples (of {tot_s}; where warps wait):")
for a in sorted(agg, key=lambda a: -a[1])[:top]:
    print(f"  {a[1] / tot_s * 100:5.1f}% samples  {a[0] / tot * 100:5.1f}% inst  {a[2]}:{a[3]:<4d} {a[4]}")
