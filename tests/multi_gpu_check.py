"""Multi-GPU parity check (run under torchrun on >= 2 GPUs of one box; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

The sharded anneal (particles split over the ranks, incremental works all-gathered on a side stream, resampled rows pulled over
NVLink inside the propagate kernel behind a peer barrier) must be BIT-IDENTICAL to the single-GPU anneal of the same global
ensemble: noise is keyed by the global particle id and every rank reduces the same weight vector with exact integer sums."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coddiwomple_b200 import testsystems as ts  # noqa: E402
from coddiwomple_b200.engine import LangevinParameters, SMCEngine, TwistSequence  # noqa: E402
from coddiwomple_b200.parallel import ShardedSMCEngine  # noqa: E402
from coddiwomple_b200.system import load_xml  # noqa: E402


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("CDW_WATCHDOG_S", "600")), exit=True)   # a hang must not hold the GPU node
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, world = dist.get_rank(), dist.get_world_size()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    xml = load_xml(os.path.join(root, "tests", "golden", "systems", "fluorobenzene.vacuum.xml.gz"))
    frames = np.load(os.path.join(root, "tests", "golden", "fluorobenzene_cache.npy"))
    ok = True
    cases = (("multinomial", 0.9999, 4096, 8, "baoab"), ("systematic", 0.9999, 4096, 8, "baoab"), ("multinomial", 0.5, 2048 * world, 12, "baoab"),
             ("multinomial", 0.9999, 4096, 8, "ula"), ("multinomial", 0.9999, 4096, 8, "twisted"), ("multinomial", 0.9999, 4096, 8, "twisted-dense"),
             ("systematic", 0.9999, 1024 * 12 * world, 6, "baoab"))   # (ula: BASELINE configs[3], the Euler-Maruyama proposal, sharded; last: a batch whose per-rank share is launched very differently from the whole)
    for kind, thr, P, T, proposal in cases:
        x0 = frames[np.random.RandomState(0).randint(0, len(frames), P)]
        seq = ts.relative_alchemical_sequence(T)
        twist = None
        if proposal == "twisted":   # controlled SMC: the Euler-Maruyama proposal with a per-iteration quadratic twist (troubleshooting/csmc.py:460-623)
            r = np.random.RandomState(9)
            twist = TwistSequence(30.0 * r.rand(T - 1, 13, 3), 2.0 * r.randn(T - 1, 13, 3), c=r.randn(T - 1))
            proposal = "ula"
        if proposal == "twisted-dense":   # the same with a full 39 x 39 matrix A_t
            r = np.random.RandomState(10)
            G = r.randn(T - 1, 39, 39) / np.sqrt(39.0)
            twist = TwistSequence.dense(10.0 * G @ np.swapaxes(G, 1, 2), 2.0 * r.randn(T - 1, 39), c=r.randn(T - 1))
            proposal = "ula"
        lang = LangevinParameters(timestep=0.0005) if proposal == "ula" else LangevinParameters()
        sharded = ShardedSMCEngine(xml, P, seq, langevin=lang, resampler=kind, floor_threshold=thr, seed=3, resample_seed=4, proposal=proposal, twist=twist)
        sharded.anneal(x0[sharded.first:sharded.last])
        if os.environ.get("CDW_CHECK_GRAPH") and proposal == "baoab" and kind == "multinomial" and thr != 0.5:
            eager_cum = sharded.cum.clone()
            sharded.capture(x0[sharded.first:sharded.last])
            sharded.anneal(x0[sharded.first:sharded.last])
            if rank == 0:
                print(f"[multi_gpu_check] CUDA-graph replay of the sharded loop == eager: {torch.equal(eager_cum, sharded.cum)}", flush=True)
            ok &= torch.equal(eager_cum, sharded.cum)
        pos_parts = [torch.empty_like(sharded.local("pos")) for _ in range(world)]
        dist.all_gather(pos_parts, sharded.local("pos").contiguous())
        vel_parts = [torch.empty_like(sharded.local("vel")) for _ in range(world)]
        dist.all_gather(vel_parts, sharded.local("vel").contiguous())
        lab_parts = [torch.empty_like(sharded.local("label")) for _ in range(world)]
        dist.all_gather(lab_parts, sharded.local("label").contiguous())
        if rank == 0:
            single = SMCEngine(xml, P, seq, langevin=lang, resampler=kind, floor_threshold=thr, seed=3, resample_seed=4, proposal=proposal, twist=twist)
            single.anneal(x0)
            e = single.ensemble
            same = (torch.equal(torch.cat(pos_parts), e.pos[e.cur]) and torch.equal(torch.cat(vel_parts), e.vel[e.cur]) and
                    torch.equal(torch.cat(lab_parts), e.labels) and torch.equal(sharded.cum, e.cum) and torch.equal(sharded.stats, single.stats))
            print(f"[multi_gpu_check] world={world} {proposal}{(' + dense twist' if twist.is_dense else ' + twist') if twist is not None else ''} {kind} thr={thr} P={P}: sharded == single-GPU bitwise: {same}; "
                  f"resampled at {sharded.resampled_iterations()} FE={sharded.free_energy():.6f}", flush=True)
            ok &= bool(same) and len(sharded.resampled_iterations()) > 0 or (thr == 0.5 and bool(same))
        sharded.close()
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
