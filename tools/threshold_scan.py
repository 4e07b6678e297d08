"""How many of the 99 iterations of the bench workload resample, per floor threshold (one anneal each)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from coddiwomple_b200.engine import SMCEngine, LangevinParameters
xml, x0, seq = bench.workload_inputs(100000, 0)
for thr in (0.5, 0.9, 0.99, 0.999):
    eng = SMCEngine(xml, 100000, seq, langevin=LangevinParameters(), seed=0, resample_seed=1, floor_threshold=thr)
    eng.anneal(x0)
    s = eng.ness_trace()
    print(f"threshold {thr}: resampled at {len(eng.resampled_iterations())} iterations; FE {eng.free_energy():.4f}; min nESS {s.min():.4f}", flush=True)
