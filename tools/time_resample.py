"""Per-call time of cdw_smc_resample (+ the sharded loop's weight_update) replayed from a CUDA graph, as the anneal uses it."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from coddiwomple_b200 import _lib  # noqa: E402

lib = _lib.load()
for n in [int(a) for a in sys.argv[1:]] or [10000, 20000, 40000, 80000]:
    g = torch.Generator("cuda").manual_seed(5)
    W = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    Wc = torch.empty_like(W)
    cum = torch.zeros_like(W)
    key = torch.full((1,), -1, dtype=torch.int64, device="cuda")
    stats = torch.zeros(8, dtype=torch.float64, device="cuda")
    anc = torch.zeros(n, dtype=torch.int32, device="cuda")
    cdf = torch.zeros(n, dtype=torch.int64, device="cuda")
    u = torch.zeros(n, dtype=torch.float64, device="cuda")
    wsb = int(lib.cdw_weight_workspace_bytes(n))
    ws = torch.zeros((wsb + 7) // 8, dtype=torch.int64, device="cuda")

    def body(reps):
        st = _lib.current_stream()
        for k in range(reps):
            key.fill_(-1)
            _lib.check(lib.cdw_weight_update(_lib.ptr(W), None, None, None, n, None, _lib.ptr(Wc), _lib.ptr(key), st))
            _lib.check(lib.cdw_smc_resample(0, _lib.ptr(W), n, -1.0, _lib.ptr(key), C.c_uint64(1), C.c_uint64(k), _lib.ptr(stats), _lib.ptr(cum), _lib.ptr(cum),
                                            _lib.ptr(anc), _lib.ptr(cdf), _lib.ptr(u), _lib.ptr(ws), wsb, st))

    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        body(3)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        body(50)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    print(f"n={n}: fill + weight_update + cdw_smc_resample = {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per call (graph replay)")
