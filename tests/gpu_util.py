"""Helpers for the -m gpu tests: thin wrappers that call the C ABI through the product's ctypes binding."""
import ctypes as C

import numpy as np
import torch

from coddiwomple_b200 import _lib
from coddiwomple_b200.engine import DeviceSystem, positions_to_rows


def rows(x):
    return positions_to_rows(np.asarray(x, dtype=np.float32), "cuda")


def lam_dev(spec, lam):
    return torch.from_numpy(spec.lambda_vector(lam)).cuda()


def energy_forces(ds, x, lam):
    lib = _lib.load()
    r = rows(x)
    P = r.shape[0]
    e = torch.empty(P, dtype=torch.float64, device="cuda")
    f = torch.empty_like(r)
    _lib.check(lib.cdw_energy_forces(ds.handle, _lib.ptr(r), P, _lib.ptr(lam_dev(ds.spec, lam)), _lib.ptr(e), _lib.ptr(f), _lib.current_stream()), "cdw_energy_forces")
    return e.cpu().numpy(), f[..., :3].double().cpu().numpy()


def reduced_potential(ds, x, lam, beta):
    lib = _lib.load()
    r = rows(x)
    P = r.shape[0]
    u = torch.empty(P, dtype=torch.float64, device="cuda")
    _lib.check(lib.cdw_reduced_potential(ds.handle, _lib.ptr(r), P, _lib.ptr(lam_dev(ds.spec, lam)), float(beta), _lib.ptr(u), _lib.current_stream()),
               "cdw_reduced_potential")
    return u.cpu().numpy()


def smc_propagate(ds, x, v, lam, dt, gamma, kT, n_steps, seed, step0, gid0=0, tol=1e-6, flags=0, anc=None, aux=None, cum=None, labels=None, splitting=0,
                  twist=None):
    """Same contract as tests.host_harness.HostSystem.smc_propagate, on the GPU."""
    lib = _lib.load()
    pos = rows(x)
    vel = rows(v) if v is not None else None
    P = pos.shape[0]

    def t(a, dt_):
        return None if a is None else torch.as_tensor(np.ascontiguousarray(a), dtype=dt_).cuda()

    anc_d, aux_d, cum_d, lab_d = t(anc, torch.int32), t(aux, torch.float64), t(cum, torch.float64), t(labels, torch.int32)
    o = dict(pos=torch.zeros_like(pos), vel=torch.zeros_like(pos))
    for k in ("inc", "w", "aux", "u_first", "u_last", "shadow", "proposal"):
        o[k] = torch.zeros(P, dtype=torch.float64, device="cuda")
    o["label"] = torch.zeros(P, dtype=torch.int32, device="cuda")
    o["nan"] = torch.zeros(P, dtype=torch.int32, device="cuda")
    o["wmin"] = torch.full((1,), -1, dtype=torch.int64, device="cuda")
    trials = torch.zeros(P, 2, dtype=torch.int32, device="cuda")
    dense = None
    if isinstance(twist, dict):   # dense twist: theta, chol, A, b, beta (numpy), logk0, cd
        keep = {k: torch.as_tensor(np.ascontiguousarray(twist[k]), dtype=torch.float64).cuda() for k in ("theta", "chol", "A", "b", "beta")}
        keep["scratch"] = torch.zeros(P, 2 * keep["b"].numel(), dtype=torch.float64, device="cuda")
        dense = _lib.DenseTwist(keep["theta"].data_ptr(), keep["chol"].data_ptr(), keep["A"].data_ptr(), keep["b"].data_ptr(), keep["beta"].data_ptr(),
                                float(twist["logk0"]), float(twist["cd"]), keep["scratch"].data_ptr())
        twist = None
    twist_d = None if twist is None else torch.as_tensor(np.ascontiguousarray(twist), dtype=torch.float64).cuda()
    prm = _lib.LangevinParams(dt, gamma, kT, tol, n_steps, flags, int(splitting), trials.data_ptr() if splitting else None,
                              None if twist_d is None else twist_d.data_ptr(), None if dense is None else C.addressof(dense))
    _lib.check(lib.cdw_smc_propagate(ds.handle, _lib.ptr(pos), _lib.ptr(vel), _lib.ptr(o["pos"]), _lib.ptr(o["vel"]), P, _lib.ptr(anc_d), _lib.ptr(aux_d),
                                     _lib.ptr(cum_d), _lib.ptr(lab_d), _lib.ptr(lam_dev(ds.spec, lam)), C.byref(prm), C.c_uint64(seed), C.c_uint64(step0),
                                     C.c_int64(gid0), _lib.ptr(o["inc"]), _lib.ptr(o["w"]), _lib.ptr(o["aux"]), _lib.ptr(o["label"]), _lib.ptr(o["u_first"]),
                                     _lib.ptr(o["u_last"]), _lib.ptr(o["shadow"]), _lib.ptr(o["proposal"]), _lib.ptr(o["nan"]), _lib.ptr(o["wmin"]),
                                     _lib.current_stream()), "cdw_smc_propagate")
    out = {k: v_.cpu().numpy() for k, v_ in o.items()}
    out["x"] = out["pos"][..., :3].astype(np.float64)
    out["v"] = out["vel"][..., :3].astype(np.float64)
    out["wmin"] = out["wmin"].view(np.uint64)
    out["trials"] = trials.cpu().numpy()
    return out


def smc_lookahead(ds, x, v, f, lam, lam_next, dt, gamma, kT, n_steps, seed, step0, gid0=0, tol=1e-6, flags=0, anc=None, labels=None):
    """Same contract as tests.host_harness.HostSystem.smc_lookahead, on the GPU (cdw_smc_lookahead)."""
    lib = _lib.load()
    pos = rows(x)
    vel = rows(v) if v is not None else None
    frc = None if f is None else torch.as_tensor(np.ascontiguousarray(f, dtype=np.float32)).cuda()
    P = pos.shape[0]
    anc_d = None if anc is None else torch.as_tensor(np.ascontiguousarray(anc), dtype=torch.int32).cuda()
    lab_d = None if labels is None else torch.as_tensor(np.ascontiguousarray(labels), dtype=torch.int32).cuda()
    o = dict(pos=torch.zeros_like(pos), vel=torch.zeros_like(pos), force=torch.zeros_like(pos), inc_next=torch.zeros(P, dtype=torch.float64, device="cuda"),
             aux=torch.zeros(P, dtype=torch.float64, device="cuda"), label=torch.zeros(P, dtype=torch.int32, device="cuda"),
             nan=torch.zeros(P, dtype=torch.int32, device="cuda"))
    prm = _lib.LangevinParams(dt, gamma, kT, tol, n_steps, flags)
    lam_t = lam_dev(ds.spec, lam)                                         # (named: a temporary would be recycled by the allocator
    lam_n = lam_dev(ds.spec, lam_next) if lam_next is not None else None  #  before the launch and the two rows would alias)
    _lib.check(lib.cdw_smc_lookahead(ds.handle, _lib.ptr(pos), _lib.ptr(vel), _lib.ptr(frc), _lib.ptr(o["pos"]) if n_steps else None,
                                     _lib.ptr(o["vel"]) if n_steps else None, _lib.ptr(o["force"]), P, _lib.ptr(anc_d), _lib.ptr(lab_d), _lib.ptr(lam_t),
                                     _lib.ptr(lam_n), C.byref(prm), C.c_uint64(seed), C.c_uint64(step0),
                                     C.c_int64(gid0), _lib.ptr(o["inc_next"]) if lam_next is not None else None, _lib.ptr(o["aux"]), _lib.ptr(o["label"]),
                                     _lib.ptr(o["nan"]), _lib.current_stream()), "cdw_smc_lookahead")
    out = {k: v_.cpu().numpy() for k, v_ in o.items()}
    out["x"] = out["pos"][..., :3].astype(np.float64)
    out["v"] = out["vel"][..., :3].astype(np.float64)
    return out
