"""Short driver for ncu: a few SMC iterations of the bench workload (BASELINE.json configs[1]) on one GPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from coddiwomple_b200 import testsystems as ts  # noqa: E402
from coddiwomple_b200.engine import LangevinParameters, SMCEngine  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
xml, _ = ts.alanine_dipeptide_shaped_xml()
seq = ts.relative_alchemical_sequence(100)
eng = SMCEngine(xml, P, seq, langevin=LangevinParameters(), seed=0, resample_seed=1, always_resample=("--always" in sys.argv))
eng.initialize(ts.alanine_dipeptide_shaped_ensemble(P))
for t in range(1, iters + 1):
    eng.step(t)
torch.cuda.synchronize()
print("ok", eng.free_energy())
