"""Per-launch table of an .ncu-rep holding several kernels (the resampling path): duration, DRAM bytes read / written, achieved DRAM GB/s, the
share of the measured HBM copy peak, L2 hit rate, issue-slot use.  Needs the ncu CLI only (no GPU).
    python tools/ncu_kernel_table.py gpurun_out/<dir>/resample.ncu-rep [hbm_peak_gbs] > profiles/<name>.txt"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
if peak is None:
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6551.7
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
ix = {k: i for i, k in enumerate(hdr)}


def val(r, k, scale_to=None):
    if k not in ix or r[ix[k]] in ("", "n/a"):
        return float("nan")
    v = float(r[ix[k]].replace(",", ""))
    u = units[ix[k]]
    if scale_to == "B":
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    if scale_to == "us":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    return v


print(f"# {os.path.basename(rep)}: ncu --set full --clock-control none, one row per profiled launch; HBM peak = {peak:.1f} GB/s (measured copy bandwidth)")
print(f"{'kernel':34s} {'grid':>7s} {'blk':>4s} {'regs':>4s} {'us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>7s} {'of peak':>7s} {'L2 hit%':>7s} {'issue%':>6s} {'warps%':>6s}")
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("cdw::", "")
    us = val(r, "gpu__time_duration.sum", "us")
    rd, wr = val(r, "dram__bytes_read.sum", "B"), val(r, "dram__bytes_write.sum", "B")
    gbs = (rd + wr) / (us * 1e-6) / 1e9
    print(f"{name[:34]:34s} {int(val(r, 'launch__grid_size')):7d} {int(val(r, 'launch__block_size')):4d} {int(val(r, 'launch__registers_per_thread')):4d} {us:8.1f} "
          f"{rd / 1e6:8.2f} {wr / 1e6:8.2f} {gbs:7.0f} {gbs / peak:7.3f} {val(r, 'lts__t_sector_hit_rate.pct'):7.1f} "
          f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} {val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f}")
