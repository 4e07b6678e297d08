"""Aggregate an .ncu-rep's source page by enclosing function (instruction and sample shares)."""
import csv, re, subprocess, sys, os
from collections import defaultdict
rep = sys.argv[1]
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "coddiwomple_b200", "csrc")
starts = {}
for fn in os.listdir(root):
    marks = []
    for i, l in enumerate(open(os.path.join(root, fn)), 1):
        m = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:CDW_HD|__global__|__device__|static|inline|CDW_DEV)\b.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", l)
        if m and not l.startswith(" "):
            marks.append((i, m.group(1)))
    starts[fn] = marks
def func(fn, line):
    name = "?"
    for i, n in starts.get(fn, []):
        if i <= line: name = n
        else: break
    return name
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hd = None, None
agg = defaultdict(lambda: [0, 0])
for r in csv.reader(src.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        cur, hd = r[1].split("/")[-1], None; continue
    if r and r[0] == "Line No":
        hd = r; continue
    if hd is None or len(r) < len(hd): continue
    d = dict(zip(hd, r))
    try:
        k = (cur, func(cur, int(d["Line No"])))
        agg[k][0] += int(d["Instructions Executed"] or 0); agg[k][1] += int(d["# Samples"] or 0)
    except (ValueError, KeyError):
        continue
ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k[0]:22s} {k[1]:28s} inst={v[0]/max(ti,1):6.3f} smp={v[1]/max(ts,1):6.3f}")
