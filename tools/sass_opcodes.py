"""Opcode histogram of the kernels in libcdw_b200.so (cuobjdump -sass): the evidence for what the binary runs on —
UBLKCP (TMA bulk copy of the table images), LDGSTS (cp.async row gathers), DFMA/DMUL/DADD (fp64 pipe), F2F (fp32 <-> fp64), MUFU.
    python tools/sass_opcodes.py > profiles/sass_opcodes_r2.txt"""
import os
import re
import subprocess
import sys
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "coddiwomple_b200", "libcdw_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
per = defaultdict(Counter)
arch = set()
name = None
for l in sass.splitlines():
    m = re.search(r"arch = (sm_\w+)", l)
    if m:
        arch.add(m.group(1))
    m = re.search(r"Function : (\S+)", l)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", l)
    if m and name:
        per[name][m.group(1).split(".")[0]] += 1
print(f"# {os.path.relpath(so, ROOT)}: cubins for {sorted(arch)}; static SASS instruction counts per kernel (opcode families)")
keys = ["UBLKCP", "LDGSTS", "LDG", "STG", "LDS", "STS", "DFMA", "DMUL", "DADD", "FFMA", "FMUL", "FADD", "F2F", "I2F", "F2I", "MUFU", "SHFL", "BAR", "ATOM", "ATOMG", "RED",
        "SYNCS", "HMMA", "UTCHMMA"]
print(f"{'kernel':58s} {'total':>6s} " + " ".join(f"{k:>6s}" for k in keys))
for n in sorted(per, key=lambda n: -sum(per[n].values())):
    c = per[n]
    print(f"{n[:58]:58s} {sum(c.values()):6d} " + " ".join(f"{c.get(k, 0):6d}" for k in keys))
