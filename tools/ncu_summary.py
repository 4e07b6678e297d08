"""Summarise an .ncu-rep: key raw metrics and the hottest source lines (needs the ncu CLI; no GPU)."""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, vals = rows[0], rows[2] if len(rows) > 2 else rows[1]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_pipe_lsu.sum",
        "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_xu.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
for h, v in zip(hdr, vals):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.08):
        print(f"{h:85s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hd = None, None
agg = defaultdict(lambda: [0, 0, 0])
text = {}
for r in csv.reader(src.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        cur, hd = r[1].split("/")[-1], None
        continue
    if r and r[0] == "Line No":
        hd = r
        continue
    if hd is None or len(r) < len(hd):
        continue
    d = dict(zip(hd, r))
    try:
        key = (cur, int(d["Line No"]))
        agg[key][0] += int(d["Instructions Executed"] or 0); agg[key][1] += int(d["# Samples"] or 0); agg[key][2] += int(d["Thread Instructions Executed"] or 0)
        text[key] = r[1][:100]
    except (ValueError, KeyError):
        continue
ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print(f"total warp-inst {ti}  samples {ts}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0]:16s}:{k[1]:4d} inst={v[0]/max(ti,1):6.3f} smp={v[1]/max(ts,1):6.3f} act={v[2]/max(v[0],1):5.1f} | {text[k]}")
