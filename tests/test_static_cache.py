"""Static-term cache (cdw_static_cache): the lambda-independent part of the force field is evaluated once per
coordinate set and handed from the closing evaluation of one SMC iteration to the opening evaluation of the next.
CPU: the device code (host harness) with the cache against the same code without it and against the oracle."""
import numpy as np
import pytest

from coddiwomple_b200 import _lib, testsystems as ts
from coddiwomple_b200.system import spec_from_xml
from oracle import langevin
from oracle.potential import XmlSystemOracle, kT_kj_per_mol
from tests import host_harness as hh
from tests.helpers import cache_frames, hybrid_positions, relative_lambdas, system_xml

kT = kT_kj_per_mol()


def _pair(xml):
    return XmlSystemOracle(xml), spec_from_xml(xml), hh.HostSystem(spec_from_xml(xml))


def test_static_split_counts():
    """Perses vacuum hybrids: every bonded term and the environment pairs are static; the synthetic alanine system
    scales a 10-atom subset."""
    for name, expect_bonded_static in (("fluorobenzene", True), ("ethylbenzene", True)):
        hs = hh.HostSystem(spec_from_xml(system_xml(name)))
        c = hs.spec.counts()
        n_static = hs.info()["n_slots"]
        assert n_static >= c["bonds"] + c["angles"] + c["torsions"]
        assert n_static < c["bonds"] + c["angles"] + c["torsions"] + c["pairs"]   # some pairs are alchemical


@pytest.mark.parametrize("name,flags,n_steps", [("fluorobenzene", 1, 1), ("fluorobenzene", 0, 2), ("ethylbenzene", 1, 1), ("fluorobenzene", _lib.CDW_FLAG_ULA, 2)])
def test_cached_iterations_equal_uncached(name, flags, n_steps):
    o, spec, hs = _pair(system_xml(name))
    X = cache_frames()[:6].astype(np.float64) if name == "fluorobenzene" else np.repeat(hybrid_positions(name)[None], 6, 0).astype(np.float32).astype(np.float64)
    P = X.shape[0]
    V = np.zeros_like(X)
    dt = 0.0005 if flags & _lib.CDW_FLAG_ULA else 0.002
    lam0, lam1, lam2 = (spec.lambda_vector(relative_lambdas(q)) for q in (0.2, 0.3, 0.4))
    # seed: n_steps = 0 evaluates the static part at the input rows
    seed_out = hs.smc_propagate_cached(X, V, lam0, dt, 1.0, kT, 0, seed=7, step0=0)
    plain0 = hs.smc_propagate(X, V, lam0, dt, 1.0, kT, 0, seed=7, step0=0)
    np.testing.assert_allclose(seed_out["u_first"], plain0["u_first"], rtol=1e-13)
    U_static = o.energy_forces(X, relative_lambdas(0.2))  # sanity: the cached energy is lambda-independent
    seed_b = hs.smc_propagate_cached(X, V, lam2, dt, 1.0, kT, 0, seed=7, step0=0)
    np.testing.assert_array_equal(seed_out["cache"][1], seed_b["cache"][1])
    np.testing.assert_array_equal(seed_out["cache"][0], seed_b["cache"][0])
    # iteration 1 and 2, with resampling in between, cached vs uncached
    a1 = hs.smc_propagate_cached(X, V, lam1, dt, 1.0, kT, n_steps, seed=7, step0=1, flags=flags, aux=-seed_out["u_first"], cum=np.zeros(P), cache=seed_out["cache"])
    b1 = hs.smc_propagate(X, V, lam1, dt, 1.0, kT, n_steps, seed=7, step0=1, flags=flags, aux=-plain0["u_first"], cum=np.zeros(P))
    for k, tol in (("x", 2e-7), ("v", 2e-5), ("u_first", 1e-10), ("u_last", 2e-4), ("inc", 5e-4), ("proposal", 5e-4)):
        assert np.abs(a1[k] - b1[k]).max() <= tol, k
    anc = np.array([2, 2, 0, 5, 4, 4])
    a2 = hs.smc_propagate_cached(a1["x"], a1["v"], lam2, dt, 1.0, kT, n_steps, seed=7, step0=1 + n_steps, flags=flags & ~1, anc=anc, aux=a1["aux"], cum=a1["w"],
                                 cache=a1["cache"])
    b2 = hs.smc_propagate(a1["x"], a1["v"], lam2, dt, 1.0, kT, n_steps, seed=7, step0=1 + n_steps, flags=flags & ~1, anc=anc, aux=a1["aux"], cum=a1["w"])
    for k, tol in (("x", 2e-7), ("v", 2e-5), ("u_first", 1e-10), ("u_last", 2e-4), ("inc", 5e-4)):
        assert np.abs(a2[k] - b2[k]).max() <= tol, k
    # the cache written at the output rows is what a fresh seeding at those rows gives
    fresh = hs.smc_propagate_cached(a2["x"], a2["v"], lam0, dt, 1.0, kT, 0, seed=7, step0=0)
    np.testing.assert_allclose(a2["cache"][1], fresh["cache"][1], rtol=1e-13)
    np.testing.assert_allclose(a2["cache"][0], fresh["cache"][0], rtol=1e-6, atol=1e-3)


def test_cached_step_matches_oracle():
    o, spec, hs = _pair(system_xml("fluorobenzene"))
    X = cache_frames()[:5].astype(np.float64)
    P = X.shape[0]
    lam = relative_lambdas(0.3)
    V = np.random.RandomState(1).normal(0, 0.3, X.shape).astype(np.float32).astype(np.float64)
    langevin._rattle_v(X, V, 1.0 / o.masses, o.constraints, 1e-6)
    V = V.astype(np.float32).astype(np.float64)
    ref = langevin.baoab(lambda xx: o.energy_forces(xx, lam), X, V, o.masses, o.constraints, 0.002, 1.0, kT, 2, seed=7, gids=np.arange(5, 5 + P), step0=11)
    seed_out = hs.smc_propagate_cached(X, V, spec.lambda_vector(lam), 0.002, 1.0, kT, 0, seed=7, step0=0)
    out = hs.smc_propagate_cached(X, V, spec.lambda_vector(lam), 0.002, 1.0, kT, 2, seed=7, step0=11, gid0=5, aux=np.full(P, 0.25), cum=np.full(P, 1.5),
                                  cache=seed_out["cache"])
    assert np.abs(out["x"] - ref["x"]).max() < 2e-7
    assert np.abs(out["v"] - ref["v"]).max() < 2e-5
    np.testing.assert_allclose(out["u_first"], ref["u_first"] / kT, rtol=1e-12)
    np.testing.assert_allclose(out["u_last"], ref["u_last"] / kT, atol=2e-4)


def test_nan_guard_keeps_the_ancestors_cache():
    o, spec, hs = _pair(system_xml("fluorobenzene"))
    X = cache_frames()[:2].astype(np.float64)
    X[1, 12] = X[1, 5]
    V = np.zeros_like(X)
    lam = spec.lambda_vector(relative_lambdas(0.2))
    seed_out = hs.smc_propagate_cached(X, V, lam, 0.002, 1.0, kT, 0, seed=1, step0=0)
    out = hs.smc_propagate_cached(X, V, lam, 0.002, 1.0, kT, 1, seed=1, step0=1, aux=np.zeros(2), cum=np.zeros(2), cache=seed_out["cache"])
    assert list(out["nan"]) == [0, 1]
    np.testing.assert_array_equal(out["x"][1], X[1])
    np.testing.assert_array_equal(out["cache"][0][1], seed_out["cache"][0][1])
    assert out["cache"][1][1] == seed_out["cache"][1][1] or (np.isnan(out["cache"][1][1]) and np.isnan(seed_out["cache"][1][1]))


@pytest.mark.parametrize("flags", [0, 1])
def test_split_launch_equals_one_launch(flags):
    """CDW_FLAG_PHASE_OPEN followed by CDW_FLAG_PHASE_REST (the sharded loop's way to overlap the weight exchange with the
    integration) gives bit for bit what the unsplit call gives."""
    o, spec, hs = _pair(system_xml("fluorobenzene"))
    X = cache_frames()[:6].astype(np.float64)
    P = X.shape[0]
    V = np.random.RandomState(2).normal(0, 0.2, X.shape).astype(np.float32).astype(np.float64)
    lam0, lam1 = (spec.lambda_vector(relative_lambdas(q)) for q in (0.2, 0.3))
    seed_out = hs.smc_propagate_cached(X, V, lam0, 0.002, 1.0, kT, 0, seed=7, step0=0)
    anc = np.array([1, 1, 0, 5, 4, 2])
    kw = dict(anc=anc, aux=-seed_out["u_first"], cum=np.arange(P) * 0.1, cache=seed_out["cache"])
    full = hs.smc_propagate_cached(X, V, lam1, 0.002, 1.0, kT, 2, seed=7, step0=3, flags=flags, **kw)
    first = hs.smc_propagate_cached(X, V, lam1, 0.002, 1.0, kT, 2, seed=7, step0=3, flags=flags | _lib.CDW_FLAG_PHASE_OPEN, **kw)
    np.testing.assert_array_equal(first["inc"], full["inc"])
    np.testing.assert_array_equal(first["w"], full["w"])
    rest = hs.smc_propagate_cached(X, V, lam1, 0.002, 1.0, kT, 2, seed=7, step0=3, flags=flags | _lib.CDW_FLAG_PHASE_REST, handover=first["handover"], **kw)
    for k in ("x", "v", "aux", "u_last", "nan"):
        np.testing.assert_array_equal(rest[k], full[k])
    np.testing.assert_array_equal(rest["cache"][0], full["cache"][0])
    np.testing.assert_array_equal(rest["cache"][1], full["cache"][1])
