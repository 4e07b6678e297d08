"""Resampling-only sweep (BASELINE.json config 5): K2 statistics, K3a integer-cdf scan + search, K3b gather, 1e3 .. 1e8
particles; log-weights ~ N(0, s^2); GB/s against SURVEY.md 8(d)'s algorithmic bytes.  `--once N` runs each kernel once
at N particles (for an ncu launch list)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from coddiwomple_b200 import _lib, kernels  # noqa: E402

lib = _lib.load()
st = _lib.current_stream()
if "--l2-fetch" in sys.argv:   # experiment: cudaLimitMaxL2FetchGranularity (32 / 64 / 128 bytes) for the random-row gathers
    import ctypes
    import glob
    torch.zeros(1, device="cuda")
    rt = ctypes.CDLL(glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))[0])
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(sys.argv[sys.argv.index("--l2-fetch") + 1])))
    got = ctypes.c_size_t(0)
    rt.cudaDeviceGetLimit(ctypes.byref(got), 5)
    print(f"# cudaLimitMaxL2FetchGranularity: rc {rc}, now {got.value} bytes", file=sys.stderr)


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def run(n, s, A=13, reps=5, with_vel=True):
    g = torch.Generator("cuda").manual_seed(5)
    W = torch.randn(n, dtype=torch.float64, device="cuda", generator=g) * s
    u = kernels.philox_uniforms(6, 0, n)
    ws = kernels.WeightWorkspace(n)
    anc = torch.empty(n, dtype=torch.int32, device="cuda")
    Wc = torch.empty_like(W)

    def stats():
        ws.wmin_key.fill_(-1)
        _lib.check(lib.cdw_weight_update(_lib.ptr(W), None, None, None, n, None, _lib.ptr(Wc), _lib.ptr(ws.wmin_key), st))
        _lib.check(lib.cdw_weight_stats(_lib.ptr(W), n, -1.0, _lib.ptr(ws.wmin_key), _lib.ptr(ws.stats), _lib.ptr(ws.buf), ws.bytes, st))

    def resample(kind):
        _lib.check(lib.cdw_resample_indices(kind, _lib.ptr(W), _lib.ptr(u), n, _lib.ptr(ws.stats), None, None, _lib.ptr(anc), _lib.ptr(ws.cdf), _lib.ptr(ws.buf), ws.bytes, st))

    t_stats = timed(stats, reps)
    ness = float(ws.stats[_lib.STAT_NESS])
    t_sys = timed(lambda: resample(1), reps)
    t_mul = timed(lambda: resample(0), reps)
    anc_m = anc.clone()
    pos = torch.randn(n, A, 4, device="cuda")
    vel = torch.randn(n, A, 4, device="cuda") if with_vel else None
    pos_o = torch.empty_like(pos)
    vel_o = torch.empty_like(pos) if with_vel else None
    aux = torch.randn(n, dtype=torch.float64, device="cuda"); aux_o = torch.empty_like(aux)
    lab = torch.arange(n, dtype=torch.int32, device="cuda"); lab_o = torch.empty_like(lab)

    def gather(a):
        _lib.check(lib.cdw_gather_particles(_lib.ptr(pos), _lib.ptr(vel), _lib.ptr(a), _lib.ptr(pos_o), _lib.ptr(vel_o), _lib.ptr(aux), _lib.ptr(aux_o),
                                            _lib.ptr(lab), _lib.ptr(lab_o), n, A, st))

    t_gm = timed(lambda: gather(anc_m), reps)
    resample(1)
    t_gs = timed(lambda: gather(anc), reps)
    rows = 2 if with_vel else 1
    gb = lambda b, ms: b / (ms * 1e-3) / 1e9
    return dict(n=n, s=s, nESS=ness, ms=dict(stats=t_stats, systematic=t_sys, multinomial=t_mul, gather_multinomial=t_gm, gather_systematic=t_gs),
                gbs=dict(stats=gb(n * 24.0, t_stats), scan_search_systematic=gb(n * 28.0, t_sys), scan_search_multinomial=gb(n * 36.0, t_mul),
                         gather_multinomial=gb(n * (2 * rows * A * 16.0 + 28), t_gm), gather_systematic=gb(n * (2 * rows * A * 16.0 + 28), t_gs)),
                particles_per_s=dict(multinomial=n / ((t_stats + t_mul + t_gm) * 1e-3), systematic=n / ((t_stats + t_sys + t_gs) * 1e-3)))


if __name__ == "__main__":
    if "--once" in sys.argv:
        n = int(sys.argv[sys.argv.index("--once") + 1])
        print(json.dumps(run(n, 1.0, reps=1)))
    else:
        out = []
        for n in (10 ** 3, 10 ** 4, 10 ** 5, 10 ** 6, 10 ** 7, 10 ** 8):
            for s in ((0.1, 1.0, 3.0) if n <= 10 ** 7 else (1.0,)):
                r = run(n, s, with_vel=(n <= 10 ** 7), reps=5 if n <= 10 ** 7 else 3)
                out.append(r)
                print(json.dumps(r), flush=True)
