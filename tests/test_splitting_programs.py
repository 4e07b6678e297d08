"""Programmable OMMLI splittings (cdw_langevin_params.splitting), including the Metropolised block "{ ... }" of
coddiwomple/openmm/integrators.py:427-444: the device code (host harness) against the oracle interpreter and against the fixed
"V R O R V" body."""
import numpy as np
import pytest

from coddiwomple_b200 import testsystems as ts
from coddiwomple_b200.system import spec_from_xml
from oracle import langevin
from oracle.potential import XmlSystemOracle, kT_kj_per_mol
from tests import host_harness as hh
from tests.helpers import cache_frames, relative_lambdas, system_xml

kT = kT_kj_per_mol()


def _setup(P=5):
    xml = system_xml("fluorobenzene")
    spec = spec_from_xml(xml)
    hs = hh.HostSystem(spec)
    X = cache_frames()[:P].astype(np.float64)
    # the cache frames are PDB coordinates (3 decimals): two steps put them on the constraint manifold, where a rejected move returns to
    X = hs.smc_propagate(X, np.zeros_like(X), spec.lambda_vector(relative_lambdas(0.3)), 0.002, 1.0, kT, 2, seed=1, step0=1, flags=1)["x"]
    return XmlSystemOracle(xml), spec, hs, X


def test_encode_splitting():
    assert langevin.encode_splitting("V R O R V") == 0x12321
    assert langevin.encode_splitting("O { V R V } O") == 0x3512143


@pytest.mark.parametrize("splitting,dt", [("V R O R V", 0.002), ("O V R V O", 0.002), ("O { V R V } O", 0.004), ("{ V R O R V }", 0.006), ("R V O V R", 0.002)])
def test_program_matches_the_oracle_interpreter(splitting, dt):
    o, spec, hs, X = _setup()
    P = X.shape[0]
    lam = relative_lambdas(0.3)
    V = np.zeros_like(X)
    ref = langevin.langevin_program(lambda xx: o.energy_forces(xx, lam), X, V, o.masses, o.constraints, dt, 5.0, kT, 3, seed=7, gids=np.arange(5, 5 + P),
                                    step0=11, splitting=splitting, randomize_velocities=True)
    for team in (1, 4):
        hh.set_team(team)
        out = hs.smc_propagate(X, V, spec.lambda_vector(lam), dt, 5.0, kT, 3, seed=7, step0=11, gid0=5, flags=1, aux=np.zeros(P), cum=np.zeros(P),
                               splitting=langevin.encode_splitting(splitting))
        hh.set_team(1)
        np.testing.assert_array_equal(out["trials"][:, 0], ref["trials"])
        np.testing.assert_array_equal(out["trials"][:, 1], ref["accepted"])
        assert np.abs(out["x"] - ref["x"]).max() < 5e-7 and np.abs(out["v"] - ref["v"]).max() < 5e-5
        np.testing.assert_allclose(out["u_last"], ref["u_last"] / kT, atol=5e-4)
        np.testing.assert_allclose(out["shadow"], ref["shadow_work"] / kT, atol=1e-3)
        np.testing.assert_allclose(out["proposal"], ref["proposal_work"] / kT, atol=1e-4)
        for (i, j, d) in o.constraints:
            assert np.abs(np.linalg.norm(out["x"][:, i] - out["x"][:, j], axis=-1) / d - 1).max() < 3e-6


def test_default_program_equals_the_fixed_body():
    """The program 'V R O R V' walks the same trajectory as the fixed BAOAB body with work tracking (same noise counters)."""
    o, spec, hs, X = _setup()
    P = X.shape[0]
    lamv = spec.lambda_vector(relative_lambdas(0.3))
    V = np.zeros_like(X)
    a = hs.smc_propagate(X, V, lamv, 0.002, 1.0, kT, 2, seed=3, step0=4, flags=3, aux=np.zeros(P), cum=np.zeros(P))
    b = hs.smc_propagate(X, V, lamv, 0.002, 1.0, kT, 2, seed=3, step0=4, flags=1, aux=np.zeros(P), cum=np.zeros(P), splitting=langevin.encode_splitting("V R O R V"))
    for k in ("x", "v", "inc", "u_first", "u_last", "aux"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    np.testing.assert_allclose(a["shadow"], b["shadow"], atol=1e-9)
    np.testing.assert_allclose(a["proposal"], b["proposal"], atol=1e-12)


def test_metropolised_block_rejects_and_restores():
    """A step that is far too long: (almost) every trial is rejected, the positions come back and the velocities are flipped."""
    o, spec, hs, X = _setup(P=4)
    P = X.shape[0]
    lamv = spec.lambda_vector(relative_lambdas(0.3))
    V = np.random.RandomState(0).normal(0, 0.3, X.shape)
    langevin._rattle_v(X, V, 1.0 / o.masses, o.constraints, 1e-6)
    V = V.astype(np.float32).astype(np.float64)
    out = hs.smc_propagate(X, V, lamv, 0.05, 1.0, kT, 1, seed=3, step0=4, aux=np.zeros(P), cum=np.zeros(P), splitting=langevin.encode_splitting("{ V R V }"))
    rejected = out["trials"][:, 1] == 0
    assert rejected.any() and (out["trials"][:, 0] == 1).all()
    np.testing.assert_array_equal(out["x"][rejected], X[rejected])
    np.testing.assert_array_equal(out["v"][rejected], -V[rejected])
    np.testing.assert_array_equal(out["shadow"], np.zeros(P))                                  # reset at the end of the block
    np.testing.assert_allclose(out["u_last"][rejected], out["u_first"][rejected], rtol=1e-13)  # the energy of the restored state
