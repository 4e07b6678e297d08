"""Latency of the sharded loop's collective: all_gather_into_tensor of P_loc doubles per rank, replayed from a CUDA graph."""
import os, sys
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
world, rank = dist.get_world_size(), dist.get_rank()
for p_loc in (10000,):
    src = torch.randn(p_loc, dtype=torch.float64, device="cuda")
    dst = torch.empty(p_loc * world, dtype=torch.float64, device="cuda")
    tok = torch.zeros(1, device="cuda")
    for name, fn in (("all_gather 80KB/rank", lambda: dist.all_gather_into_tensor(dst, src)), ("all_reduce 4B", lambda: dist.all_reduce(tok))):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(5):
                fn()
        torch.cuda.synchronize(); dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            for _ in range(50):
                fn()
        g.replay(); torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        if rank == 0:
            print(f"world={world} NCCL_PROTO={os.environ.get('NCCL_PROTO','-')} NCCL_ALGO={os.environ.get('NCCL_ALGO','-')} {name}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per call", flush=True)
dist.destroy_process_group()
