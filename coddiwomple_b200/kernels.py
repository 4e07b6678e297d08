"""Thin, array-level entry points to the K2/K3 kernels for callers that hold plain arrays (numpy or torch).

Everything here runs on the GPU through libcdw_b200.so; numpy inputs are staged to the device and results are
returned in the caller's flavour.  No CPU arithmetic path exists.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _dev(a, dtype=torch.float64):
    if a is None:
        return None
    if torch.is_tensor(a):
        return a.to(device="cuda", dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


class WeightWorkspace:
    def __init__(self, n):
        self.lib = _lib.load()
        self.n = int(n)
        self.bytes = int(self.lib.cdw_weight_workspace_bytes(self.n))
        self.buf = torch.zeros((self.bytes + 7) // 8, dtype=torch.int64, device="cuda")
        self.wmin_key = torch.zeros(1, dtype=torch.int64, device="cuda")
        self.stats = torch.zeros(_lib.STATS_LEN, dtype=torch.float64, device="cuda")
        self.cdf = torch.zeros(self.n, dtype=torch.int64, device="cuda")


def weight_stats(works, floor_threshold=0.5, always=False, incremental=None, workspace=None):
    """Returns (W device tensor, stats device tensor[8], workspace).  W = works (+ incremental)."""
    lib = _lib.load()
    cum = _dev(works)
    inc = _dev(incremental)
    n = cum.numel()
    ws = workspace or WeightWorkspace(n)
    ws.wmin_key.fill_(-1)
    W = torch.empty_like(cum)
    zero = torch.zeros_like(cum) if inc is None else inc
    _lib.check(lib.cdw_weight_update(_lib.ptr(zero), None, None, _lib.ptr(cum), n, None, _lib.ptr(W), _lib.ptr(ws.wmin_key), _lib.current_stream()),
               "cdw_weight_update")
    _lib.check(lib.cdw_weight_stats(_lib.ptr(W), n, -1.0 if always else float(floor_threshold), _lib.ptr(ws.wmin_key), _lib.ptr(ws.stats),
                                    _lib.ptr(ws.buf), ws.bytes, _lib.current_stream()), "cdw_weight_stats")
    return W, ws.stats, ws


def device_ness(cumulative_works, incremental_works=None):
    _, stats, _ = weight_stats(cumulative_works, incremental=incremental_works, always=True)
    return float(stats[_lib.STAT_NESS].item())


def device_normalized_weights(works):
    W, stats, _ = weight_stats(works, always=True)
    s = stats.cpu().numpy()
    w = torch.exp(-(W - s[_lib.STAT_WMIN])) / s[_lib.STAT_SUM_E]
    return w if torch.is_tensor(works) else w.cpu().numpy()


def resample_indices(works, uniforms, kind="multinomial", workspace=None):
    """Ancestors for cumulative works `works` (always resamples): K2 stats + K3a integer cdf + search.

    uniforms: N doubles (multinomial) or one double u0 (systematic).  Returns int32 ancestors (same flavour as input).
    """
    lib = _lib.load()
    W, stats, ws = weight_stats(works, always=True, workspace=workspace)
    n = W.numel()
    u = _dev(np.atleast_1d(uniforms) if not torch.is_tensor(uniforms) else uniforms)
    k = _lib.CDW_RESAMPLE_SYSTEMATIC if kind == "systematic" else _lib.CDW_RESAMPLE_MULTINOMIAL
    if k == _lib.CDW_RESAMPLE_MULTINOMIAL and u.numel() != n:
        raise ValueError("multinomial resampling needs one uniform per particle")
    anc = torch.empty(n, dtype=torch.int32, device="cuda")
    _lib.check(lib.cdw_resample_indices(k, _lib.ptr(W), _lib.ptr(u), n, _lib.ptr(stats), None, None, _lib.ptr(anc), _lib.ptr(ws.cdf), _lib.ptr(ws.buf),
                                        ws.bytes, _lib.current_stream()), "cdw_resample_indices")
    return anc if torch.is_tensor(works) else anc.cpu().numpy()


def philox_uniforms(seed, stream_id, n):
    lib = _lib.load()
    out = torch.empty(int(n), dtype=torch.float64, device="cuda")
    _lib.check(lib.cdw_philox_uniforms(C.c_uint64(seed), C.c_uint64(stream_id), int(n), _lib.ptr(out), _lib.current_stream()), "cdw_philox_uniforms")
    return out


def gather_particles(pos, vel, anc, aux=None, label=None):
    """K3b on float4-row tensors: returns the gathered copies."""
    lib = _lib.load()
    P, A = pos.shape[0], pos.shape[1]
    pos_o = torch.empty_like(pos)
    vel_o = torch.empty_like(vel) if vel is not None else None
    aux_o = torch.empty_like(aux) if aux is not None else None
    lab_o = torch.empty_like(label) if label is not None else None
    _lib.check(lib.cdw_gather_particles(_lib.ptr(pos), _lib.ptr(vel), _lib.ptr(anc), _lib.ptr(pos_o), _lib.ptr(vel_o), _lib.ptr(aux), _lib.ptr(aux_o),
                                        _lib.ptr(label), _lib.ptr(lab_o), P, A, _lib.current_stream()), "cdw_gather_particles")
    return pos_o, vel_o, aux_o, lab_o


def fma_peak(fp64=True, iters=4096):
    lib = _lib.load()
    t, ms = C.c_double(), C.c_double()
    _lib.check(lib.cdw_fma_peak(1 if fp64 else 0, int(iters), C.byref(t), C.byref(ms)), "cdw_fma_peak")
    return t.value, ms.value
