"""Static SASS instruction count of one kernel, attributed to the enclosing source function (needs -lineinfo):
   python tools/sass_by_function.py <obj-or-so> <kernel-name-substring>"""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

obj, pat = sys.argv[1], sys.argv[2]
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "coddiwomple_b200", "csrc")
starts = {}
for fn in os.listdir(root):
    marks = []
    for i, l in enumerate(open(os.path.join(root, fn)), 1):
        m = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:CDW_HD|CDW_NOINLINE_HD|__global__|__device__|static|inline)\b.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", l)
        if m and not l.startswith(" "):
            marks.append((i, m.group(1)))
    starts[fn] = marks


def func(fn, line):
    name = "?"
    for i, n in starts.get(fn, []):
        if i <= line:
            name = n
        else:
            break
    return name


with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
    total = Counter()
    for cubin in os.listdir(d):
        sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
        inside, cur = False, ("?", 0)
        for l in sass.splitlines():
            if l.startswith(".text."):
                inside = pat in l
                continue
            if not inside:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l):
                total[(cur[0], func(cur[0], cur[1]))] += 1
        if total:
            break
n = sum(total.values())
print(f"{n} SASS instructions in kernels matching {pat!r}")
for k, v in total.most_common(30):
    print(f"{v:6d} {v / n:6.3f}  {k[0]}:{k[1]}")
