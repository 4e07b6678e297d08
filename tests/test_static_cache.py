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


def _two_iterations(hs, spec, X, V, n_steps, lookahead, bad_row=None):
    """Two SMC iterations (lambda 0.2 -> 0.3 -> 0.4, a resampling gather in between) through the cached path or the look-ahead
    path of the device code; returns what both must agree on."""
    P = X.shape[0]
    lam0, lam1, lam2 = (spec.lambda_vector(relative_lambdas(q)) for q in (0.2, 0.3, 0.4))
    anc = np.array([1, 1, 0, 5, 4, 2])[:P]
    if not lookahead:
        seed = hs.smc_propagate_cached(X, V, lam0, 0.002, 1.0, kT, 0, seed=7, step0=0)
        a1 = hs.smc_propagate_cached(X, V, lam1, 0.002, 1.0, kT, n_steps, seed=7, step0=n_steps, flags=1, aux=-seed["u_first"], cum=np.zeros(P), cache=seed["cache"])
        a2 = hs.smc_propagate_cached(a1["x"], a1["v"], lam2, 0.002, 1.0, kT, n_steps, seed=7, step0=2 * n_steps, anc=anc, aux=a1["aux"], cum=np.zeros(P),
                                     cache=a1["cache"])
        return dict(aux0=-seed["u_first"], inc1=a1["inc"], x1=a1["x"], v1=a1["v"], aux1=a1["aux"], inc2=a2["inc"], x2=a2["x"], v2=a2["v"], aux2=a2["aux"],
                    nan1=a1["nan"], nan2=a2["nan"])
    e = hs.smc_lookahead(X, V, None, lam0, lam1, 0.002, 1.0, kT, 0, seed=7, step0=0)
    b1 = hs.smc_lookahead(X, V, e["force"], lam1, lam2, 0.002, 1.0, kT, n_steps, seed=7, step0=n_steps, flags=1)
    b2 = hs.smc_lookahead(b1["x"], b1["v"], b1["force"], lam2, None, 0.002, 1.0, kT, n_steps, seed=7, step0=2 * n_steps, anc=anc, labels=np.arange(P) + 10)
    assert list(b2["label"]) == list(anc + 10)
    return dict(aux0=e["aux"], inc1=e["inc_next"], x1=b1["x"], v1=b1["v"], aux1=b1["aux"], inc2=b1["inc_next"][anc], x2=b2["x"], v2=b2["v"], aux2=b2["aux"],
                nan1=b1["nan"], nan2=b2["nan"])


@pytest.mark.parametrize("n_steps", [1, 3])
def test_lookahead_iterations_equal_cached_iterations(n_steps):
    """cdw_smc_lookahead moves the evaluation that opens iteration t + 1 to the end of launch t (the next weights become a
    by-product, the forces travel with the particle): states, auxiliary works and incremental works are BITWISE those of the
    cached path, where launch t + 1 opens with that evaluation."""
    o, spec, hs = _pair(system_xml("fluorobenzene"))
    X = cache_frames()[:6].astype(np.float64)
    V = np.zeros_like(X)
    a = _two_iterations(hs, spec, X, V, n_steps, lookahead=False)
    b = _two_iterations(hs, spec, X, V, n_steps, lookahead=True)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    # and the oracle: incremental work of iteration 1 is u_1(x_0) - u_0(x_0)
    np.testing.assert_allclose(b["inc1"], o.reduced_potential(X, relative_lambdas(0.3)) - o.reduced_potential(X, relative_lambdas(0.2)), rtol=1e-8, atol=1e-10)


def test_lookahead_reverts_a_move_that_ends_in_a_non_finite_energy():
    """openmm/propagators.py:166-186: the replica keeps its pre-move state; its aux and its next incremental work are those of
    the state that is kept (what openmm/coddiwomple.py:315, 326-327 compute on the reverted particle)."""
    o, spec, hs = _pair(system_xml("fluorobenzene"))
    X = cache_frames()[:4].astype(np.float64)
    V = np.zeros_like(X)
    V[2, 0, 0] = np.inf   # finite at rest, not after a step
    lam1, lam2 = (spec.lambda_vector(relative_lambdas(q)) for q in (0.3, 0.4))
    e = hs.smc_lookahead(X, V, None, lam1, lam1, 0.002, 1.0, kT, 0, seed=7, step0=0)
    for team in (1, 4):
        hh.set_team(team)
        b = hs.smc_lookahead(X, V, e["force"], lam1, lam2, 0.002, 1.0, kT, 1, seed=7, step0=1)
        hh.set_team(1)
        assert list(b["nan"]) == [0, 0, 1, 0]
        np.testing.assert_array_equal(b["pos"][2], hs.rows(X)[2])
        np.testing.assert_array_equal(b["vel"][2], hs.rows(V)[2])
        keep = hs.smc_lookahead(X[2:3], V[2:3], None, lam1, lam2, 0.002, 1.0, kT, 0, seed=7, step0=0)
        assert np.isfinite(keep["aux"][0]) and np.isfinite(keep["inc_next"][0])
        np.testing.assert_allclose(b["aux"][2], keep["aux"][0], rtol=1e-13)              # (team size changes the summation order)
        np.testing.assert_allclose(b["inc_next"][2], keep["inc_next"][0], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(b["aux"][2], -o.reduced_potential(X[2:3], relative_lambdas(0.3))[0], rtol=1e-10)
        # the other replicas are untouched by their neighbour's revert
        ok = hs.smc_lookahead(X[[0, 1]], V[[0, 1]], e["force"][[0, 1]], lam1, lam2, 0.002, 1.0, kT, 1, seed=7, step0=1)
        if team == 1:
            np.testing.assert_array_equal(ok["pos"], b["pos"][:2])
            np.testing.assert_array_equal(ok["inc_next"], b["inc_next"][:2])
