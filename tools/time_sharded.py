"""Time the sharded anneal of the bench workload (1e5 fluorobenzene-hybrid particles per GPU, 100 lambda) under torchrun, for A/B runs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/time_sharded.py [--never] [--label X]
Prints ms per anneal (CUDA events, max over ranks, mean of 5 after 3 warm-ups; CUDA-graphed loop)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from coddiwomple_b200 import testsystems as ts  # noqa: E402
from coddiwomple_b200.engine import LangevinParameters, positions_to_rows  # noqa: E402
from coddiwomple_b200.parallel import ShardedSMCEngine  # noqa: E402
from coddiwomple_b200.system import load_xml  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rank, world = dist.get_rank(), dist.get_world_size()
P = int(os.environ.get("CDW_P", "100000"))
golden = os.path.join(ROOT, "tests", "golden")
xml = load_xml(os.path.join(golden, "systems", "fluorobenzene.vacuum.xml.gz"))
frames = np.load(os.path.join(golden, "fluorobenzene_cache.npy"))
never = "--never" in sys.argv
eng = ShardedSMCEngine(xml, P * world, ts.relative_alchemical_sequence(100), langevin=LangevinParameters(), seed=0, resample_seed=1,
                       floor_threshold=0.5, always_resample=not never)
rows = positions_to_rows(frames[np.random.RandomState(3 + rank).randint(0, len(frames), P)]).cuda()
eng.capture(rows)
for _ in range(3):
    eng.anneal(rows)
ms = []
for _ in range(5):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.anneal(rows); e1.record(); e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms.append(float(t.item()))
if rank == 0:
    label = sys.argv[sys.argv.index("--label") + 1] if "--label" in sys.argv else ""
    print(f"[time_sharded] world={world} P/GPU={P} {'never' if never else 'always'} resampling {label}: {np.mean(ms):.2f} ms per anneal "
          f"({1e3 * np.mean(ms) / 99:.1f} us per iteration), FE={eng.free_energy():.4f}", flush=True)
eng.close()
dist.destroy_process_group()
