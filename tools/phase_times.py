"""Where does the propagate kernel spend its time?  Times K1, E+F (n_steps = 0), 1 and 2 BAOAB steps on the bench workload."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from coddiwomple_b200 import _lib, testsystems as ts  # noqa: E402
from coddiwomple_b200.engine import DeviceSystem, LangevinParameters, positions_to_rows  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
xml, _ = ts.alanine_dipeptide_shaped_xml()
ds = DeviceSystem(xml)
lib = _lib.load()
pos = positions_to_rows(ts.alanine_dipeptide_shaped_ensemble(P))
vel = torch.zeros_like(pos)
pos2, vel2 = torch.empty_like(pos), torch.empty_like(pos)
lam = torch.from_numpy(ds.spec.lambda_vector(ts.relative_alchemical_sequence(100)[50])).cuda()
u = torch.empty(P, dtype=torch.float64, device="cuda")
f = torch.empty_like(pos)
st = _lib.current_stream()
lp = LangevinParameters()


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts_ = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts_.append(e0.elapsed_time(e1) * 1e3)
    return np.median(ts_)


def k1():
    _lib.check(lib.cdw_reduced_potential(ds.handle, _lib.ptr(pos), P, _lib.ptr(lam), 0.4, _lib.ptr(u), st))


def ef():
    _lib.check(lib.cdw_energy_forces(ds.handle, _lib.ptr(pos), P, _lib.ptr(lam), _lib.ptr(u), _lib.ptr(f), st))


def step(n, flags=0):
    prm = lp.as_struct(flags, n_steps=n)
    _lib.check(lib.cdw_smc_propagate(ds.handle, _lib.ptr(pos), _lib.ptr(vel), _lib.ptr(pos2), _lib.ptr(vel2), P, None, None, None, None, _lib.ptr(lam), C.byref(prm),
                                     C.c_uint64(1), C.c_uint64(1), C.c_int64(0), None, _lib.ptr(u), None, None, None, None, None, None, None, None, st))


t_k1, t_ef, t_0, t_1, t_2, t_1t = timeit(k1), timeit(ef), timeit(lambda: step(0)), timeit(lambda: step(1)), timeit(lambda: step(2)), timeit(lambda: step(1, 2))
print(f"P={P} team={os.environ.get('CDW_TEAM','auto')}: K1 {t_k1:.1f} us | E+F (force out) {t_ef:.1f} | step n=0 {t_0:.1f} | n=1 {t_1:.1f} | n=2 {t_2:.1f} | n=1 +works {t_1t:.1f}")
print(f"   => one E+F evaluation ~ {t_0:.1f} us; one step's integrator+constraints+noise ~ {t_2 - t_1 - t_0 if False else (t_2 - t_1) - (t_1 - t_0) + 0:.1f} (step total {t_2 - t_1:.1f} incl. one evaluation)")
