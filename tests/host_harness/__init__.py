"""ctypes wrapper of the TEST-ONLY host build of the per-replica device code (see harness.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

from coddiwomple_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libcdw_host_harness.so")
ROOT = os.path.dirname(os.path.dirname(HERE))


def build():
    src = os.path.join(HERE, "harness.cpp")
    deps = [src] + [os.path.join(ROOT, "coddiwomple_b200", "csrc", f) for f in ("cdw_math.cuh", "cdw_lane.cuh", "cdw_pack.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++20", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-o", SO, src])
    return SO


_h = None


def lib():
    global _h
    if _h is None:
        _h = C.CDLL(build())
        _h.hh_last_error.restype = C.c_char_p
        if os.environ.get("CDW_HH_TEAM"):   # run a whole test file through the team emulation
            _h.hh_set_team(int(os.environ["CDW_HH_TEAM"]))
    return _h


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def set_team(T):
    """Lanes per replica for the propagate entry points (1 = one lane per replica; > 1 = host threads emulating a warp team)."""
    lib().hh_set_team(int(T))


class HostSystem:
    def __init__(self, spec):
        self.spec = spec
        desc, self._keep = _lib.make_desc(spec)
        self.handle = C.c_void_p()
        rc = lib().hh_system_create(C.byref(desc), C.byref(self.handle))
        if rc != 0:
            raise RuntimeError(lib().hh_last_error().decode())

    def info(self):
        ns, nc, se, ss = C.c_int(), C.c_int(), C.c_size_t(), C.c_size_t()
        lib().hh_system_info(self.handle, C.byref(ns), C.byref(nc), C.byref(se), C.byref(ss))
        return dict(n_slots=ns.value, n_clusters=nc.value, smem_energy=se.value, smem_step=ss.value)

    @staticmethod
    def rows(x):
        """(P, A, 3) -> float4 rows (P, A, 4) float32."""
        x = np.asarray(x)
        out = np.zeros(x.shape[:2] + (4,), dtype=np.float32)
        out[..., :3] = x
        return out

    def reduced_potential(self, x, lam, beta):
        pos = self.rows(x)
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        u = np.zeros(pos.shape[0])
        lib().hh_reduced_potential(self.handle, _p(pos), C.c_int64(pos.shape[0]), _p(lam), C.c_double(beta), _p(u))
        return u

    def energy_forces(self, x, lam):
        pos = self.rows(x)
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        e = np.zeros(pos.shape[0])
        f = np.zeros_like(pos)
        lib().hh_energy_forces(self.handle, _p(pos), C.c_int64(pos.shape[0]), _p(lam), _p(e), _p(f))
        return e, f[..., :3].astype(np.float64)

    def smc_propagate(self, x, v, lam, dt, gamma, kT, n_steps, seed, step0, gid0=0, tol=1e-6, flags=0, anc=None, aux=None, cum=None, labels=None, splitting=0,
                      twist=None):
        pos = self.rows(x)
        vel = self.rows(v) if v is not None else None
        P = pos.shape[0]
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        trials = np.zeros((P, 2), dtype=np.int32)
        dense = None
        if isinstance(twist, dict):   # dense twist: theta, chol, A, b, beta (numpy), logk0, cd
            keep = {k: np.ascontiguousarray(twist[k], dtype=np.float64) for k in ("theta", "chol", "A", "b", "beta")}
            keep["scratch"] = np.zeros((P, 2 * keep["b"].size))
            dense = _lib.DenseTwist(keep["theta"].ctypes.data, keep["chol"].ctypes.data, keep["A"].ctypes.data, keep["b"].ctypes.data,
                                    keep["beta"].ctypes.data, float(twist["logk0"]), float(twist["cd"]), keep["scratch"].ctypes.data)
            twist = None
        twist = None if twist is None else np.ascontiguousarray(twist, dtype=np.float64)
        prm = _lib.LangevinParams(dt, gamma, kT, tol, n_steps, flags, int(splitting), trials.ctypes.data if splitting else None,
                                  None if twist is None else twist.ctypes.data, None if dense is None else C.addressof(dense))
        o = dict(pos=np.zeros_like(pos), vel=np.zeros_like(pos), inc=np.zeros(P), w=np.zeros(P), aux=np.zeros(P), label=np.zeros(P, dtype=np.int32),
                 u_first=np.zeros(P), u_last=np.zeros(P), shadow=np.zeros(P), proposal=np.zeros(P), nan=np.zeros(P, dtype=np.int32),
                 wmin=np.array([2 ** 64 - 1], dtype=np.uint64))
        anc = None if anc is None else np.ascontiguousarray(anc, dtype=np.int32)
        aux = None if aux is None else np.ascontiguousarray(aux, dtype=np.float64)
        cum = None if cum is None else np.ascontiguousarray(cum, dtype=np.float64)
        labels = None if labels is None else np.ascontiguousarray(labels, dtype=np.int32)
        lib().hh_smc_propagate(self.handle, _p(pos), _p(vel), _p(o["pos"]), _p(o["vel"]), C.c_int64(P), _p(anc), _p(aux), _p(cum), _p(labels), _p(lam),
                               C.byref(prm), C.c_uint64(seed), C.c_uint64(step0), C.c_int64(gid0), _p(o["inc"]), _p(o["w"]), _p(o["aux"]), _p(o["label"]),
                               _p(o["u_first"]), _p(o["u_last"]), _p(o["shadow"]), _p(o["proposal"]), _p(o["nan"]), _p(o["wmin"]))
        o["x"] = o["pos"][..., :3].astype(np.float64)
        o["v"] = o["vel"][..., :3].astype(np.float64)
        o["trials"] = trials
        return o

    def smc_lookahead(self, x, v, f, lam, lam_next, dt, gamma, kT, n_steps, seed, step0, gid0=0, tol=1e-6, flags=0, anc=None, labels=None):
        """cdw_smc_lookahead on the host.  x, v: (P, A, 3); f: (P, A, 4) float32 force rows F(x; lam) (None with n_steps = 0);
        lam_next None on the last iteration.  Returns dict(x, v, pos, vel, force, inc_next, aux, label, nan)."""
        pos = self.rows(x)
        vel = self.rows(v) if v is not None else None
        frc = None if f is None else np.ascontiguousarray(f, dtype=np.float32)
        P = pos.shape[0]
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        lam_next = None if lam_next is None else np.ascontiguousarray(lam_next, dtype=np.float64)
        prm = _lib.LangevinParams(dt, gamma, kT, tol, n_steps, flags)
        o = dict(pos=np.zeros_like(pos), vel=np.zeros_like(pos), force=np.zeros_like(pos), inc_next=np.zeros(P), aux=np.zeros(P),
                 label=np.zeros(P, dtype=np.int32), nan=np.zeros(P, dtype=np.int32))
        anc = None if anc is None else np.ascontiguousarray(anc, dtype=np.int32)
        labels = None if labels is None else np.ascontiguousarray(labels, dtype=np.int32)
        lib().hh_smc_lookahead(self.handle, _p(pos), _p(vel), _p(frc), _p(o["pos"]), _p(o["vel"]), _p(o["force"]), C.c_int64(P), _p(anc), _p(labels),
                               _p(lam), _p(lam_next), C.byref(prm), C.c_uint64(seed), C.c_uint64(step0), C.c_int64(gid0),
                               _p(o["inc_next"]) if lam_next is not None else None, _p(o["aux"]), _p(o["label"]), _p(o["nan"]))
        o["x"] = o["pos"][..., :3].astype(np.float64)
        o["v"] = o["vel"][..., :3].astype(np.float64)
        return o

    def smc_propagate_cached(self, x, v, lam, dt, gamma, kT, n_steps, seed, step0, gid0=0, tol=1e-6, flags=0, anc=None, aux=None, cum=None, cache=None):
        """cdw_smc_propagate_cached on the host: `cache` = (fs rows (P, A, 4) float32, us (P,)) at the input rows, or None.
        Returns the usual outputs plus out["cache"] at the output rows."""
        pos = self.rows(x)
        vel = self.rows(v) if v is not None else None
        P = pos.shape[0]
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        prm = _lib.LangevinParams(dt, gamma, kT, tol, n_steps, flags)
        o = dict(pos=np.zeros_like(pos), vel=np.zeros_like(pos), inc=np.zeros(P), w=np.zeros(P), aux=np.zeros(P), u_first=np.zeros(P), u_last=np.zeros(P),
                 proposal=np.zeros(P), nan=np.zeros(P, dtype=np.int32))
        anc = None if anc is None else np.ascontiguousarray(anc, dtype=np.int32)
        aux = None if aux is None else np.ascontiguousarray(aux, dtype=np.float64)
        cum = None if cum is None else np.ascontiguousarray(cum, dtype=np.float64)
        fs_in = None if cache is None else np.ascontiguousarray(cache[0], dtype=np.float32)
        us_in = None if cache is None else np.ascontiguousarray(cache[1], dtype=np.float64)
        fs_out, us_out = np.zeros_like(pos), np.zeros(P)
        lib().hh_smc_propagate_cached(self.handle, _p(pos), _p(vel), _p(o["pos"]), _p(o["vel"]), C.c_int64(P), _p(anc), _p(aux), _p(cum), _p(lam), C.byref(prm),
                                      C.c_uint64(seed), C.c_uint64(step0), C.c_int64(gid0), _p(o["inc"]), _p(o["w"]), _p(o["aux"]), _p(o["u_first"]),
                                      _p(o["u_last"]), _p(o["proposal"]), _p(o["nan"]), _p(fs_in), _p(us_in), _p(fs_out), _p(us_out))
        o["x"] = o["pos"][..., :3].astype(np.float64)
        o["v"] = o["vel"][..., :3].astype(np.float64)
        o["cache"] = (fs_out, us_out)
        return o

    def __del__(self):
        try:
            lib().hh_system_destroy(self.handle)
        except Exception:
            pass
