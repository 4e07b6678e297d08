"""CPU baseline "B2" (SURVEY.md §8d): the SMC anneal compiled for the host (C++ / OpenMP, all cores), see smc_port.cpp.
TEST / BENCHMARK INFRASTRUCTURE ONLY: imported by bench.py's cpu_baseline / --impl reference leg and by tests, never by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BUILD = os.path.join(os.path.dirname(HERE), "_build")
SO = os.path.join(BUILD, "libcdw_cpu_port.so")


def build(force=False):
    """g++ -O3 -march=native -fopenmp; output under oracle/_build/ (git-ignored, travels to the GPU box with the snapshot... and
    is rebuilt there if the host CPU differs: -march=native code is not portable, so the library is keyed by the CPU model)."""
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(HERE, "smc_port.cpp")
    deps = [src] + [os.path.join(ROOT, "coddiwomple_b200", "csrc", f) for f in ("cdw_math.cuh", "cdw_lane.cuh", "cdw_pack.h")]
    tag = os.path.join(BUILD, "cpu_port.host")
    host = _cpu_model()
    stale = force or not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps)
    if not stale and (not os.path.exists(tag) or open(tag).read() != host):
        stale = True
    if stale:
        subprocess.check_call(["g++", "-O3", "-march=native", "-fopenmp", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", SO, src])
        with open(tag, "w") as fh:
            fh.write(host)
    return SO


def _cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for ln in fh:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


_h = None


def lib():
    global _h
    if _h is None:
        _h = C.CDLL(build())
        _h.port_last_error.restype = C.c_char_p
    return _h


def threads():
    return int(lib().port_threads())


def use_all_cores():
    """Run on every core this process may use, whatever OMP_NUM_THREADS says (torchrun sets it to 1 for every rank); returns the count."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().port_set_threads(int(n))
    return threads()


class CpuPort:
    def __init__(self, spec):
        from coddiwomple_b200 import _lib   # struct definitions of include/cdw.h only: nothing of the product runs here
        self._lib_defs = _lib
        self.spec = spec
        desc, self._keep = _lib.make_desc(spec)
        self.handle = C.c_void_p()
        if lib().port_system_create(C.byref(desc), C.byref(self.handle)) != 0:
            raise RuntimeError(lib().port_last_error().decode())

    def anneal(self, x0, parameter_sequence, dt=0.002, gamma=1.0, kT=2.4943386436406856, tol=1e-6, n_steps=1, floor_threshold=0.5, seed=0, resample_seed=1):
        """x0 (P, A, 3) nm.  Returns dict(x, cum, free_energy, nESS, resampled_at)."""
        x0 = np.asarray(x0)
        P, A = x0.shape[0], x0.shape[1]
        pos = np.zeros((P, A, 4), dtype=np.float32)
        pos[..., :3] = x0
        lam = np.ascontiguousarray(np.stack([self.spec.lambda_vector(p) for p in parameter_sequence]), dtype=np.float64)
        T = lam.shape[0]
        prm = self._lib_defs.LangevinParams(dt, gamma, kT, tol, n_steps, 0)
        cum = np.zeros(P)
        stats = np.zeros((T, 2))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = lib().port_anneal(self.handle, p(pos), C.c_int64(P), p(lam), C.c_int32(T), C.byref(prm), C.c_double(floor_threshold), C.c_uint64(seed),
                               C.c_uint64(resample_seed), p(cum), p(stats))
        if rc != 0:
            raise RuntimeError("port_anneal failed")
        m = (-cum).max()
        fe = float(-(m + np.log(np.exp(-cum - m).sum())) + np.log(P))
        return dict(x=pos[..., :3].astype(np.float64), cum=cum, free_energy=fe, nESS=stats[1:, 0], resampled_at=[t for t in range(1, T) if stats[t, 1]])

    def __del__(self):
        try:
            lib().port_system_destroy(self.handle)
        except Exception:
            pass
