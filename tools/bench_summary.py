"""One-screen summary of a bench.py JSON line (file argument or stdin)."""
import json
import sys

txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
print("N=%d value %.4g %s  ms/anneal %.2f  e2e %.2f ms  frac %.3f  resampled %s  probe %s" % (
    d["n_gpus"], d["value"], d["unit"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["resampled_iterations"], d["parity_probe"]))
for k in ("threshold_0.5", "config2", "config4", "resampling_sweep", "resampling_sweep_sharded"):
    if d.get(k):
        print(k, json.dumps(d[k])[:900])
if d.get("cpu_baseline"):
    c = d["cpu_baseline"]
    print("cpu_baseline %.4g on %d cores (%s); ratio %.1f" % (c["value"], c["cores"], c["kind"], d["e2e"]["value"] / c["value"]))
print("clocks", d["clocks"])
