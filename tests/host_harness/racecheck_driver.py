"""Child process of tests/test_host_racecheck.py: runs the TEAM emulation of the device code from a ThreadSanitizer build of
the host harness (LD_PRELOAD=libtsan.so).  Every lane of a team is a host thread and every shared-memory column is plain
memory, so a missing team barrier in cdw_lane.cuh shows up as a reported data race."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from tests import host_harness as hh  # noqa: E402

hh._h = C.CDLL(os.environ["CDW_HH_TSAN_SO"])
hh._h.hh_last_error.restype = C.c_char_p
from coddiwomple_b200 import _lib, testsystems as ts  # noqa: E402
from coddiwomple_b200.system import spec_from_xml  # noqa: E402
from tests.helpers import cache_frames, relative_lambdas, system_xml  # noqa: E402

kT = 2.4943386436406856
system, T, mode = sys.argv[1], int(sys.argv[2]), sys.argv[3]
if system == "ala":
    spec = spec_from_xml(ts.alanine_dipeptide_shaped_xml()[0])
    X = ts.alanine_dipeptide_shaped_ensemble(2, seed=1).astype(np.float32).astype(np.float64)
else:
    spec = spec_from_xml(system_xml(system))
    X = cache_frames()[:2].astype(np.float64)
hs = hh.HostSystem(spec)
V = np.zeros_like(X)
lam0, lam = spec.lambda_vector(relative_lambdas(0.2)), spec.lambda_vector(relative_lambdas(0.3))
hh.set_team(T)
if mode == "baoab":        # unsplit evaluation, shadow / proposal work bookkeeping, Maxwell-Boltzmann velocities
    hs.smc_propagate(X, V, lam, 0.002, 1.0, kT, 2, seed=7, step0=3, flags=3, aux=np.zeros(2), cum=np.zeros(2))
elif mode == "cached":     # static + dynamic passes, cache rows in and out, gather through ancestors
    s0 = hs.smc_propagate_cached(X, V, lam0, 0.002, 1.0, kT, 0, seed=7, step0=0)
    kw = dict(anc=np.array([1, 0]), aux=np.zeros(2), cum=np.zeros(2), cache=s0["cache"])
    hs.smc_propagate_cached(X, V, lam, 0.002, 1.0, kT, 2, seed=7, step0=3, flags=1, **kw)
elif mode == "lookahead":  # the look-ahead iteration: forces travel with the particle, second table image, a reverted replica
    lam2 = spec.lambda_vector(relative_lambdas(0.4))
    e = hs.smc_lookahead(X, V, None, lam0, lam, 0.002, 1.0, kT, 0, seed=7, step0=0)
    b = hs.smc_lookahead(X, V, e["force"], lam, lam2, 0.002, 1.0, kT, 2, seed=7, step0=2, flags=1, anc=np.array([1, 0]))
    Vbad = V.copy()
    Vbad[1, 0, 0] = np.inf
    hs.smc_lookahead(b["x"], Vbad, b["force"], lam2, None, 0.002, 1.0, kT, 1, seed=7, step0=4)
elif mode == "ula":        # Euler-Maruyama move with constraint projection and both kernel densities
    hs.smc_propagate(X, None, lam, 0.0005, 0.0, kT, 2, seed=7, step0=3, flags=_lib.CDW_FLAG_ULA, aux=np.zeros(2), cum=np.zeros(2))
elif mode == "twisted":    # the dense twisted move: matrix-vector rows split over the lanes, vectors through the per-particle scratch
    from oracle import twisted
    A = X.shape[1]
    r = np.random.RandomState(5)
    G = r.randn(3 * A, 3 * A) / np.sqrt(3 * A)
    tw = twisted.dense_tables(15.0 * G @ G.T, 2.0 * r.randn(3 * A), 0.37, -0.11, np.repeat(kT * 0.0005 ** 2 / np.asarray(spec.masses), 3))
    hs.smc_propagate(X, None, lam, 0.0005, 0.0, kT, 2, seed=7, step0=3, flags=_lib.CDW_FLAG_ULA, aux=np.zeros(2), cum=np.zeros(2), twist=tw)
    hs.smc_propagate(X, None, lam, 0.0005, 0.0, kT, 2, seed=7, step0=3, flags=_lib.CDW_FLAG_ULA, aux=np.zeros(2), cum=np.zeros(2),
                     twist=twisted.pack_twist(40.0 * r.rand(A, 3), 3.0 * r.randn(A, 3), c=0.37))
print("racecheck-done", system, T, mode)
