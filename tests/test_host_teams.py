"""The TEAM form of the device code on the host: the T lanes of a replica run as T host threads (real barriers, an
exchange buffer for the xor shuffles — tests/host_harness/harness.cpp), so the term partition, the per-lane force copies,
their reduction, the cluster round-robin and the team sums that a warp executes on the GPU are checked here, without one."""
import numpy as np
import pytest

from coddiwomple_b200 import _lib, testsystems as ts
from coddiwomple_b200.system import spec_from_xml
from oracle import langevin, ula
from oracle.potential import XmlSystemOracle, kT_kj_per_mol
from tests import host_harness as hh
from tests.helpers import cache_frames, hybrid_positions, relative_lambdas, system_xml

kT = kT_kj_per_mol()


@pytest.fixture(autouse=True)
def _one_lane_afterwards():
    yield
    hh.set_team(1)


def _setup(name, P=4):
    xml = system_xml(name) if name != "ala" else ts.alanine_dipeptide_shaped_xml()[0]
    o, spec = XmlSystemOracle(xml), spec_from_xml(xml)
    if name == "fluorobenzene":
        X = cache_frames()[:P].astype(np.float64)
    elif name == "ala":
        X = ts.alanine_dipeptide_shaped_ensemble(P, seed=1).astype(np.float32).astype(np.float64)
    else:
        X = np.repeat(hybrid_positions(name)[None], P, 0).astype(np.float32).astype(np.float64)
    return o, spec, hh.HostSystem(spec), X


@pytest.mark.parametrize("name,T", [("fluorobenzene", 2), ("fluorobenzene", 4), ("ethylbenzene", 4), ("ala", 4), ("ala", 8)])
def test_team_baoab_matches_oracle(name, T):
    o, spec, hs, X = _setup(name)
    P = X.shape[0]
    lam = relative_lambdas(0.3)
    V = np.random.RandomState(1).normal(0, 0.3, X.shape).astype(np.float32).astype(np.float64)
    langevin._rattle_v(X, V, 1.0 / o.masses, o.constraints, 1e-6)
    V = V.astype(np.float32).astype(np.float64)
    ref = langevin.baoab(lambda xx: o.energy_forces(xx, lam), X, V, o.masses, o.constraints, 0.002, 1.0, kT, 2, seed=7, gids=np.arange(5, 5 + P), step0=11,
                         track_works=True)
    hh.set_team(T)
    out = hs.smc_propagate(X, V, spec.lambda_vector(lam), 0.002, 1.0, kT, 2, seed=7, step0=11, gid0=5, flags=2, aux=np.full(P, 0.25), cum=np.full(P, 1.5))
    assert np.abs(out["x"] - ref["x"]).max() < 2e-7
    assert np.abs(out["v"] - ref["v"]).max() < 2e-5
    np.testing.assert_allclose(out["u_first"], ref["u_first"] / kT, rtol=1e-12)
    np.testing.assert_allclose(out["u_last"], ref["u_last"] / kT, atol=2e-4)
    np.testing.assert_allclose(out["inc"], ref["u_first"] / kT + 0.25, rtol=1e-12)
    np.testing.assert_allclose(out["shadow"], ref["shadow_work"] / kT, atol=3e-4)
    np.testing.assert_allclose(out["proposal"], ref["proposal_work"] / kT, atol=2e-5)
    for (i, j, d) in o.constraints:
        assert np.abs(np.linalg.norm(out["x"][:, i] - out["x"][:, j], axis=-1) / d - 1).max() < 3e-6


@pytest.mark.parametrize("T", [4])
def test_team_static_cache_and_lookahead(T):
    """Cached iterations (static + dynamic passes, cache rows through the ancestor) and the look-ahead form with a team."""
    o, spec, hs, X = _setup("fluorobenzene", P=6)
    P = X.shape[0]
    V = np.zeros_like(X)
    lam0, lam1 = (spec.lambda_vector(relative_lambdas(q)) for q in (0.2, 0.3))
    hh.set_team(T)
    seed_out = hs.smc_propagate_cached(X, V, lam0, 0.002, 1.0, kT, 0, seed=7, step0=0)
    anc = np.array([2, 2, 0, 5, 4, 4])
    kw = dict(anc=anc, aux=-seed_out["u_first"], cum=np.zeros(P), cache=seed_out["cache"])
    cached = hs.smc_propagate_cached(X, V, lam1, 0.002, 1.0, kT, 1, seed=7, step0=1, flags=1, **kw)
    plain = hs.smc_propagate(X, V, lam1, 0.002, 1.0, kT, 1, seed=7, step0=1, flags=1, anc=anc, aux=-seed_out["u_first"], cum=np.zeros(P))
    for k, tol in (("x", 2e-7), ("v", 2e-5), ("u_first", 1e-10), ("u_last", 2e-4), ("inc", 1e-9)):
        assert np.abs(cached[k] - plain[k]).max() <= tol, k
    # the look-ahead form of the same iteration (the evaluation that opens it was made by the previous call): bitwise equal
    lam2 = spec.lambda_vector(relative_lambdas(0.4))
    eq = hs.smc_lookahead(X, V, None, lam0, lam1, 0.002, 1.0, kT, 0, seed=7, step0=0)
    np.testing.assert_array_equal(eq["aux"], -seed_out["u_first"])
    ahead = hs.smc_lookahead(X, V, eq["force"], lam1, lam2, 0.002, 1.0, kT, 1, seed=7, step0=1, flags=1, anc=anc)
    np.testing.assert_array_equal(eq["inc_next"][anc], cached["inc"])
    for k in ("x", "v", "aux"):
        np.testing.assert_array_equal(ahead[k], cached[k])
    nxt = hs.smc_propagate_cached(cached["x"], cached["v"], lam2, 0.002, 1.0, kT, 0, seed=7, step0=2, aux=cached["aux"], cum=np.zeros(P), cache=cached["cache"])
    np.testing.assert_array_equal(ahead["inc_next"], nxt["inc"])
    # the oracle, too
    ref = langevin.baoab(lambda xx: o.energy_forces(xx, relative_lambdas(0.3)), X[anc], V[anc], o.masses, o.constraints, 0.002, 1.0, kT, 1, seed=7,
                         gids=np.arange(P), step0=1, randomize_velocities=True)
    assert np.abs(cached["x"] - ref["x"]).max() < 2e-7


def test_team_ula_matches_oracle():
    o, spec, hs, X = _setup("fluorobenzene", P=5)
    P = X.shape[0]
    lam = relative_lambdas(0.3)
    ref = ula.ula(lambda xx: o.energy_forces(xx, lam), X, o.masses, o.constraints, 0.0005, kT, 2, seed=7, gids=np.arange(5, 5 + P), step0=11)
    hh.set_team(4)
    out = hs.smc_propagate(X, None, spec.lambda_vector(lam), 0.0005, 0.0, kT, 2, seed=7, step0=11, gid0=5, flags=_lib.CDW_FLAG_ULA, aux=np.full(P, 0.25),
                           cum=np.full(P, 1.5))
    assert np.abs(out["x"] - ref["x"]).max() < 2e-7
    np.testing.assert_allclose(out["proposal"], ref["proposal_work"], atol=2e-4)
    np.testing.assert_allclose(out["inc"], out["u_last"] + 0.25 + out["proposal"], rtol=1e-12)


def test_team_sizes_agree_with_each_other():
    """Only the summation order changes with the team size: one fp32 ulp in the coordinates, 1e-14 in the energies."""
    o, spec, hs, X = _setup("ala", P=3)
    V = np.zeros_like(X)
    lam = spec.lambda_vector(relative_lambdas(0.4))
    outs = {}
    for T in (1, 2, 4, 8):
        hh.set_team(T)
        outs[T] = hs.smc_propagate(X, V, lam, 0.002, 1.0, kT, 1, seed=3, step0=2, flags=1, aux=np.zeros(3), cum=np.zeros(3))
    for T in (2, 4, 8):
        assert np.abs(outs[T]["x"] - outs[1]["x"]).max() < 2e-7
        np.testing.assert_allclose(outs[T]["u_first"], outs[1]["u_first"], rtol=1e-13)
        np.testing.assert_allclose(outs[T]["u_last"], outs[1]["u_last"], atol=2e-4)


def test_whole_anneal_on_the_host_tracks_the_oracle():
    """The full loop semantics without a GPU: device lane code (team of 4, static-term cache carried from iteration to
    iteration, gather through the ancestors) + the oracle's resampling step, against oracle.smc.anneal on the same seeds."""
    from oracle import rng, smc as oracle_smc, weights
    o, spec, hs, _ = _setup("fluorobenzene", P=1)
    P = 48
    frames = cache_frames()
    x0 = frames[np.random.RandomState(1).randint(0, len(frames), P)].astype(np.float64)
    seq = ts.relative_alchemical_sequence(6)
    trace = []
    ref = oracle_smc.anneal(o, x0, seq, seed=5, resample_seed=6, trace=trace, floor_threshold=0.9)
    hh.set_team(4)
    X, V = x0.copy(), np.zeros_like(x0)
    seed_out = hs.smc_propagate_cached(X, V, spec.lambda_vector(seq[0]), 0.002, 1.0, kT, 0, seed=5, step0=0)
    aux, cum, cache, anc = -seed_out["u_first"], np.zeros(P), seed_out["cache"], None
    labels = np.arange(P)
    for t in range(1, len(seq)):
        out = hs.smc_propagate_cached(X, V, spec.lambda_vector(seq[t]), 0.002, 1.0, kT, 1, seed=5, step0=t, flags=1 if t == 1 else 0, anc=anc,
                                      aux=aux, cum=cum, cache=cache)
        if anc is not None:
            labels = labels[anc]
        step = weights.resample_step(cum, out["inc"], out["aux"], labels, uniforms=rng.uniforms53(6, P, stream=t), kind="multinomial", floor_threshold=0.9)
        if t == 1:
            np.testing.assert_allclose(out["inc"], trace[0]["inc"], rtol=1e-10)
            np.testing.assert_array_equal(step["ancestors"], trace[0]["ancestors"])
        X, V, aux, cache = out["x"], out["v"], out["aux"], out["cache"]
        cum, anc = step["cum"], (step["ancestors"] if step["resampled"] else None)
        # aux travels with the particle: resample_step already returned it gathered; the rows are gathered by the next call
        aux = out["aux"]
    fe = weights.free_energy(cum)
    assert abs(fe - ref["free_energy"]) < 0.1, (fe, ref["free_energy"])


@pytest.mark.parametrize("name,T,dense", [("fluorobenzene", 2, False), ("fluorobenzene", 4, True), ("fluorobenzene", 2, True), ("ala", 4, True)])
def test_team_twisted_euler_maruyama_matches_oracle(name, T, dense):
    """The twisted Euler-Maruyama move with the lanes of a team as threads: the rows of the dense matrix-vector products are split over the
    lanes and the vectors travel through the per-particle scratch between team barriers (a missing barrier is a wrong result here, and a data
    race under tests/test_host_racecheck.py)."""
    from oracle import twisted
    o, spec, hs, X = _setup(name)
    P, A = X.shape[0], X.shape[1]
    lam = relative_lambdas(0.3)
    dt = 0.0005
    r = np.random.RandomState(5)
    if dense:
        G = r.randn(3 * A, 3 * A) / np.sqrt(3 * A)
        s = np.repeat(kT * dt * dt / np.asarray(o.masses), 3)
        tw = twisted.dense_tables(15.0 * G @ G.T, 2.0 * r.randn(3 * A), 0.37, -0.11, s)
        ref = twisted.twisted_ula_dense(lambda xx: o.energy_forces(xx, lam), X, o.masses, o.constraints, dt, kT, 2, seed=7, gids=np.arange(5, 5 + P),
                                        step0=11, tables=tw)
    else:
        tw = twisted.pack_twist(40.0 * r.rand(A, 3), 3.0 * r.randn(A, 3), c=0.37, d=-0.11)
        ref = twisted.twisted_ula(lambda xx: o.energy_forces(xx, lam), X, o.masses, o.constraints, dt, kT, 2, seed=7, gids=np.arange(5, 5 + P), step0=11,
                                  twist=tw)
    hh.set_team(T)
    out = hs.smc_propagate(X, None, spec.lambda_vector(lam), dt, 0.0, kT, 2, seed=7, step0=11, gid0=5, flags=_lib.CDW_FLAG_ULA, aux=np.full(P, 0.25),
                           cum=np.full(P, 1.5), twist=tw)
    assert np.abs(out["x"] - ref["x"]).max() < 2e-7
    np.testing.assert_allclose(out["proposal"], ref["proposal_work"], rtol=1e-9, atol=3e-4)
    np.testing.assert_allclose(out["inc"], out["u_last"] + 0.25 + out["proposal"], rtol=1e-12)
