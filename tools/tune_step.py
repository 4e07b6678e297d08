"""Time the fused propagate kernel and a whole SMC iteration for the library selected by CDW_LIB (tuning variants)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from coddiwomple_b200 import testsystems as ts  # noqa: E402
from coddiwomple_b200.engine import LangevinParameters, SMCEngine  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
system = sys.argv[2] if len(sys.argv) > 2 else "ala"
if system == "ala":
    xml, _ = ts.alanine_dipeptide_shaped_xml()
    x0 = ts.alanine_dipeptide_shaped_ensemble(P)
else:
    from coddiwomple_b200.system import load_xml
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    xml = load_xml(os.path.join(root, "tests", "golden", "systems", f"{system}.vacuum.xml.gz"))
    frames = np.load(os.path.join(root, "tests", "golden", "fluorobenzene_cache.npy"))
    x0 = frames[np.random.RandomState(0).randint(0, len(frames), P)]
seq = ts.relative_alchemical_sequence(100)
eng = SMCEngine(xml, P, seq, langevin=LangevinParameters(), seed=0, resample_seed=1)
eng.initialize(x0)
for t in range(1, 6):
    eng.step(t)
torch.cuda.synchronize()
ev = []
for t in range(6, 46):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); eng.propagate(t); e1.record(); eng.finish_iteration(t); e2.record()
    ev.append((e0, e1, e2))
torch.cuda.synchronize()
k = np.array([a.elapsed_time(b) for a, b, _ in ev]); it = np.array([a.elapsed_time(c) for a, _, c in ev])
print(f"{os.environ.get('CDW_LIB', 'default'):50s} P={P} {system}: propagate {k.mean()*1e3:8.1f} us (min {k.min()*1e3:8.1f})  iteration {it.mean()*1e3:8.1f} us  "
      f"-> {P / (it.mean() * 1e-3):.3e} particle-steps/s   FE {eng.free_energy():.4f}")
