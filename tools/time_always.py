"""Whole-iteration time when EVERY iteration resamples (always_resample): in-launch resampling vs two launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from coddiwomple_b200 import testsystems as ts
from coddiwomple_b200.engine import LangevinParameters, SMCEngine
P = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
kind = sys.argv[2] if len(sys.argv) > 2 else "multinomial"
xml, _ = ts.alanine_dipeptide_shaped_xml()
x0 = ts.alanine_dipeptide_shaped_ensemble(P)
seq = ts.relative_alchemical_sequence(100)
eng = SMCEngine(xml, P, seq, langevin=LangevinParameters(), seed=0, resample_seed=1, always_resample=True, resampler=kind)
eng.initialize(x0)
for t in range(1, 6):
    eng.step(t)
torch.cuda.synchronize()
ev = []
for t in range(6, 46):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.step(t); e1.record(); ev.append((e0, e1))
torch.cuda.synchronize()
it = np.array([a.elapsed_time(b) for a, b in ev])
print(f"inline={'off' if os.environ.get('CDW_NO_INLINE_RESAMPLE') else 'on'} P={P} {kind} always_resample: iteration {it.mean()*1e3:.1f} us (min {it.min()*1e3:.1f})")
