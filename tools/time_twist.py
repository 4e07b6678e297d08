"""Per-iteration time of the Euler-Maruyama SMC loop on one GPU without a twist, with a diagonal and with a full-matrix twist
(fluorobenzene hybrid, 125 000 particles = one GPU's share of BASELINE configs[3]).   python tools/time_twist.py [P]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from coddiwomple_b200 import testsystems as ts  # noqa: E402
from coddiwomple_b200.engine import LangevinParameters, SMCEngine, TwistSequence  # noqa: E402
from coddiwomple_b200.system import load_xml  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
T = 21
golden = os.path.join(ROOT, "tests", "golden")
xml = load_xml(os.path.join(golden, "systems", "fluorobenzene.vacuum.xml.gz"))
frames = np.load(os.path.join(golden, "fluorobenzene_cache.npy"))
x0 = frames[np.random.RandomState(3).randint(0, len(frames), P)]
seq = ts.relative_alchemical_sequence(T)
r = np.random.RandomState(9)
G = r.randn(T - 1, 39, 39) / np.sqrt(39.0)
twists = {"none": None, "diagonal": TwistSequence(5.0 * r.rand(T - 1, 13, 3), 0.5 * r.randn(T - 1, 13, 3)),
          "dense": TwistSequence.dense(2.0 * G @ np.swapaxes(G, 1, 2), 0.5 * r.randn(T - 1, 39))}
for name, tw in twists.items():
    eng = SMCEngine(xml, P, seq, langevin=LangevinParameters(timestep=0.0005), seed=0, resample_seed=1, proposal="ula", twist=tw)
    eng.anneal(x0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.anneal(x0); e1.record(); e1.synchronize()
    print(f"[time_twist] {name:9s} P={P}: {1e3 * e0.elapsed_time(e1) / (T - 1):.1f} us per iteration, FE={eng.free_energy():.4f}", flush=True)
