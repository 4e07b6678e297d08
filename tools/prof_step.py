"""Short driver for ncu / timing: a few SMC iterations on one GPU.
  python tools/prof_step.py [P] [iters] [--system ala|fluorobenzene] [--always] [--time]
ala = synthetic 22-atom AlanineDipeptideVacuum shape (BASELINE configs[1]); fluorobenzene = 13-atom perses vacuum hybrid
(configs[2]).  --time prints CUDA-event timings per iteration instead of being a target for ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from coddiwomple_b200 import testsystems as ts  # noqa: E402
from coddiwomple_b200.engine import LangevinParameters, SMCEngine  # noqa: E402
from coddiwomple_b200.system import load_xml  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
P = int(args[0]) if len(args) > 0 else 10000
iters = int(args[1]) if len(args) > 1 else 4
system = sys.argv[sys.argv.index("--system") + 1] if "--system" in sys.argv else "ala"
if system == "ala":
    xml, _ = ts.alanine_dipeptide_shaped_xml()
    x0 = ts.alanine_dipeptide_shaped_ensemble(P)
else:
    golden = os.path.join(ROOT, "tests", "golden")
    xml = load_xml(os.path.join(golden, "systems", f"{system}.vacuum.xml.gz"))
    frames = np.load(os.path.join(golden, "fluorobenzene_cache.npy"))
    x0 = frames[np.random.RandomState(3).randint(0, len(frames), P)]
seq = ts.relative_alchemical_sequence(100)
lookahead = "--sequential" not in sys.argv
eng = SMCEngine(xml, P, seq, langevin=LangevinParameters(), seed=0, resample_seed=1, always_resample=("--always" in sys.argv), lookahead=lookahead)
eng.initialize(x0)
evs = []
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for t in range(1, iters + 1):
    if t == 3:
        t0.record()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.step(t); e1.record()
    evs.append((e0, e1))
eng.ensemble.join()
t1.record()
torch.cuda.synchronize()
if "--time" in sys.argv:
    print("launch us:", " ".join(f"{1e3 * a.elapsed_time(b):.1f}" for a, b in evs))
    if iters > 3:
        print(f"{'look-ahead' if lookahead else 'sequential'} loop: {1e3 * t0.elapsed_time(t1) / (iters - 2):.1f} us per iteration (iterations 3..{iters}, resampling included)")
print("ok", system, P, eng.free_energy(), "resampled at", eng.resampled_iterations())
