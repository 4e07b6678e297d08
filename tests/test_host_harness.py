"""The per-replica DEVICE code (csrc/cdw_replica.cuh) compiled for the host and run lane-by-lane, against the oracle.
This is how the kernels' arithmetic is checked on the build box, which has no GPU; the same checks run against the
real kernels in test_gpu_*.py.  The harness is test-only (tests/host_harness/harness.cpp)."""
import ctypes as C

import numpy as np
import pytest

from coddiwomple_b200 import testsystems as ts
from coddiwomple_b200.system import spec_from_xml
from oracle import langevin, resampling, rng
from oracle.potential import XmlSystemOracle, kT_kj_per_mol
from tests import host_harness as hh
from tests.helpers import VACUUM_SYSTEMS, cache_frames, hybrid_positions, jittered, relative_lambdas, system_xml

kT = kT_kj_per_mol()


def _pair(xml):
    return XmlSystemOracle(xml), spec_from_xml(xml), hh.HostSystem(spec_from_xml(xml))


@pytest.mark.parametrize("name", VACUUM_SYSTEMS)
def test_energy_and_forces_all_fixtures(name):
    o, spec, hs = _pair(system_xml(name))
    X = jittered(hybrid_positions(name), 6, seed=1)
    r = np.random.RandomState(2)
    for trial in range(4):
        lam = {k: float(r.uniform(0, 1)) for k in o.global_defaults if k.startswith("lambda")}
        if trial == 0:
            lam = {k: 0.0 for k in lam}
        if trial == 1:
            lam = {k: 1.0 for k in lam}
        U, F = o.energy_forces(X, lam)
        lv = spec.lambda_vector(lam)
        e, f = hs.energy_forces(X, lv)
        u = hs.reduced_potential(X, lv, 1.0 / kT)
        np.testing.assert_allclose(e, U, rtol=1e-12)            # north star: <= 1e-6 relative; fp64 kernel gives ~1e-15
        np.testing.assert_allclose(u, U / kT, rtol=1e-12)
        assert np.abs(f - F).max() <= 1e-6 * np.abs(F).max()    # forces are accumulated in fp32


def test_reference_lambda_protocol_on_cache_frames():
    """All 151 cache frames x the 10 lambda states of the reference's SMC test."""
    o, spec, hs = _pair(system_xml("fluorobenzene"))
    X = cache_frames().astype(np.float64)
    for q in np.linspace(0, 1, 10):
        lam = relative_lambdas(float(q))
        np.testing.assert_allclose(hs.reduced_potential(X, spec.lambda_vector(lam), 1 / kT), o.reduced_potential(X, lam), rtol=1e-12)


def test_synthetic_systems():
    xml, x = ts.alanine_dipeptide_shaped_xml()
    o, spec, hs = _pair(xml)
    X = ts.alanine_dipeptide_shaped_ensemble(5).astype(np.float64)
    for lam in ts.relative_alchemical_sequence(6):
        U, F = o.energy_forces(X, lam)
        e, f = hs.energy_forces(X, spec.lambda_vector(lam))
        np.testing.assert_allclose(e, U, rtol=1e-12)
        assert np.abs(f - F).max() <= 1e-6 * np.abs(F).max()
    xmlh = ts.harmonic_oscillator_xml()
    o, spec, hs = _pair(xmlh)
    X = np.random.RandomState(0).normal(0, 0.1, (7, 1, 3)).astype(np.float32).astype(np.float64)
    for lam in ts.harmonic_parameter_sequence(4):
        U, F = o.energy_forces(X, lam)
        e, f = hs.energy_forces(X, spec.lambda_vector(lam))
        np.testing.assert_allclose(e, U, rtol=1e-13)
        np.testing.assert_allclose(f, F, rtol=1e-6, atol=1e-5)


@pytest.mark.parametrize("name,n_steps,flags", [("fluorobenzene", 1, 1), ("fluorobenzene", 3, 1), ("fluorobenzene", 2, 3), ("ethylbenzene", 1, 1),
                                                ("ethylbenzene", 2, 3), ("aniline", 2, 0)])
def test_baoab_step_matches_oracle(name, n_steps, flags):
    o, spec, hs = _pair(system_xml(name))
    X = cache_frames()[:5].astype(np.float64) if name == "fluorobenzene" else np.repeat(hybrid_positions(name)[None], 5, 0).astype(np.float32).astype(np.float64)
    P = X.shape[0]
    lam = relative_lambdas(0.3)
    V = np.random.RandomState(1).normal(0, 0.3, X.shape).astype(np.float32).astype(np.float64)
    # project the starting velocities on the constraint manifold so that both sides start from a consistent state
    langevin._rattle_v(X, V, 1.0 / o.masses, o.constraints, 1e-6)
    V = V.astype(np.float32).astype(np.float64)
    ref = langevin.baoab(lambda xx: o.energy_forces(xx, lam), X, V, o.masses, o.constraints, 0.002, 1.0, kT, n_steps, seed=7, gids=np.arange(5, 5 + P),
                         step0=11, randomize_velocities=bool(flags & 1), track_works=bool(flags & 2))
    out = hs.smc_propagate(X, V, spec.lambda_vector(lam), 0.002, 1.0, kT, n_steps, seed=7, step0=11, gid0=5, flags=flags, aux=np.full(P, 0.25), cum=np.full(P, 1.5))
    assert np.abs(out["x"] - ref["x"]).max() < 2e-7          # a few ulp of the fp32 coordinates
    assert np.abs(out["v"] - ref["v"]).max() < 2e-5
    np.testing.assert_allclose(out["u_first"], ref["u_first"] / kT, rtol=1e-12)
    np.testing.assert_allclose(out["u_last"], ref["u_last"] / kT, atol=2e-4)
    np.testing.assert_allclose(out["inc"], ref["u_first"] / kT + 0.25, rtol=1e-12)
    np.testing.assert_allclose(out["w"], 1.5 + ref["u_first"] / kT + 0.25, rtol=1e-12)
    np.testing.assert_allclose(out["aux"], -out["u_last"], rtol=0, atol=0)
    if flags & 2:
        np.testing.assert_allclose(out["shadow"], ref["shadow_work"] / kT, atol=3e-4)
        np.testing.assert_allclose(out["proposal"], ref["proposal_work"] / kT, atol=2e-5)
    for (i, j, d) in o.constraints:
        assert np.abs(np.linalg.norm(out["x"][:, i] - out["x"][:, j], axis=-1) / d - 1).max() < 3e-6
    assert out["wmin"][0] == np.uint64(min(_key(w) for w in out["w"]))


def _key(v):
    b = np.float64(v).view(np.uint64)
    return int(~b & np.uint64(0xFFFFFFFFFFFFFFFF)) if b >> np.uint64(63) else int(b | np.uint64(1 << 63))


def test_gather_on_load_and_labels():
    o, spec, hs = _pair(system_xml("fluorobenzene"))
    X = cache_frames()[:6].astype(np.float64)
    anc = np.array([3, 3, 0, 5, 5, 5])
    aux = np.arange(6) * 0.5
    labels = np.array([10, 11, 12, 13, 14, 15])
    lam = spec.lambda_vector(relative_lambdas(0.5))
    out = hs.smc_propagate(X, np.zeros_like(X), lam, 0.002, 1.0, kT, 0, seed=1, step0=0, anc=anc, aux=aux, cum=np.zeros(6), labels=labels)
    np.testing.assert_array_equal(out["x"], X[anc])
    np.testing.assert_array_equal(out["label"], labels[anc])
    u = o.reduced_potential(X[anc], relative_lambdas(0.5))
    np.testing.assert_allclose(out["inc"], u + aux[anc], rtol=1e-12)
    np.testing.assert_allclose(out["aux"], -u, rtol=1e-12)  # n_steps = 0: the closing energy is the opening one


def test_nan_guard_reverts_the_move():
    """A non-finite closing energy reverts (x, v) to the caller's state and is flagged (propagators.py:163-190)."""
    o, spec, hs = _pair(system_xml("fluorobenzene"))
    X = cache_frames()[:2].astype(np.float64)
    X[1, 12] = X[1, 5]  # F on top of its carbon: r = 0 in a bond and a Coulomb pair -> inf/NaN
    V = np.zeros_like(X)
    out = hs.smc_propagate(X, V, spec.lambda_vector(relative_lambdas(0.2)), 0.002, 1.0, kT, 1, seed=1, step0=1, aux=np.zeros(2), cum=np.zeros(2))
    assert list(out["nan"]) == [0, 1]
    np.testing.assert_array_equal(out["x"][1], X[1])
    np.testing.assert_array_equal(out["v"][1], V[1])
    assert np.isfinite(out["x"][0]).all()


def test_harmonic_oscillator_step_without_constraints():
    xml = ts.harmonic_oscillator_xml()
    o, spec, hs = _pair(xml)
    prm = ts.harmonic_oscillator_parameters()
    X = np.random.RandomState(0).normal(0, 0.1, (16, 1, 3)).astype(np.float32).astype(np.float64)
    lam = ts.harmonic_parameter_sequence(3)[1]
    ref = langevin.baoab(lambda xx: o.energy_forces(xx, lam), X, np.zeros_like(X), o.masses, [], prm["timestep"], prm["collision_rate"], kT, 4, seed=3,
                         gids=np.arange(16), step0=2, randomize_velocities=True, track_works=True)
    out = hs.smc_propagate(X, None, spec.lambda_vector(lam), prm["timestep"], prm["collision_rate"], kT, 4, seed=3, step0=2, flags=3)
    np.testing.assert_allclose(out["x"], ref["x"], atol=2e-7)
    np.testing.assert_allclose(out["v"], ref["v"], atol=5e-6)
    np.testing.assert_allclose(out["shadow"], ref["shadow_work"] / kT, atol=2e-5)


def test_det_exp_bit_identical_to_oracle():
    x = -np.abs(np.random.RandomState(3).normal(0, 50, 200000))
    x[:5] = [0.0, -1e-300, -699.99, -700.0000001, -745.2]
    got = np.zeros_like(x)
    hh.lib().hh_det_exp(x.ctypes.data_as(C.c_void_p), C.c_int64(len(x)), got.ctypes.data_as(C.c_void_p))
    np.testing.assert_array_equal(got.view(np.uint64), resampling.det_exp(x).view(np.uint64))


def test_fixed_target_bit_identical_to_oracle():
    r = np.random.RandomState(4)
    u = r.random_sample(100000)
    u[:3] = [0.0, 1.0 - 2.0 ** -53, 0.5]
    for total in (2 ** 59 + 12345, 2 ** 61 - 1, 977):
        got = np.zeros(len(u), dtype=np.uint64)
        hh.lib().hh_fixed_target(u.ctypes.data_as(C.c_void_p), C.c_int64(len(u)), C.c_uint64(total), got.ctypes.data_as(C.c_void_p))
        np.testing.assert_array_equal(got, resampling.fixed_point_targets(u, total))
        assert all(int(got[i]) == min((int(u[i] * 2 ** 53) * total) >> 53, total - 1) for i in range(200))


def test_philox_and_normals_match_oracle():
    out = (C.c_uint32 * 4)()
    hh.lib().hh_philox(C.c_uint32(0x243f6a88), C.c_uint32(0x85a308d3), C.c_uint32(0x13198a2e), C.c_uint32(0x03707344), C.c_uint32(0xa4093822),
                       C.c_uint32(0x299f31d0), out)
    assert tuple(out) == (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    z = rng.normals(99, np.array([2 ** 33 + 5]), 17, 6, rng.KIND_O_STEP)[0]
    for atom in range(6):
        o3 = (C.c_double * 3)()
        hh.lib().hh_normals(C.c_uint64(99), C.c_uint64(2 ** 33 + 5), C.c_uint32(17), C.c_uint32(atom), C.c_uint32(0), o3)
        np.testing.assert_allclose(list(o3), z[atom], atol=2e-6)


def test_ordered_key_is_monotone():
    x = np.sort(np.concatenate([np.random.RandomState(0).normal(0, 100, 1000), [0.0, -0.0, 1e-300, -1e-300, 1e300, -1e300]]))
    key = np.zeros(len(x), dtype=np.uint64)
    back = np.zeros(len(x))
    hh.lib().hh_ordered_key(x.ctypes.data_as(C.c_void_p), C.c_int64(len(x)), key.ctypes.data_as(C.c_void_p), back.ctypes.data_as(C.c_void_p))
    assert np.all(np.diff(key.astype(object)) >= 0)
    np.testing.assert_array_equal(back, x)


def test_harmonic_oscillator_equilibrium_statistics_on_the_host():
    """G6 on the build box (coddiwomple/tests/test_openmm.py:192-221 re-expressed): BAOAB on the 3-d harmonic oscillator samples
    <u> = 3/2 with std sqrt(3/2); the accumulated shadow work stays small and the heat is an O(1)-per-step quantity."""
    xml = ts.harmonic_oscillator_xml()
    o, spec, hs = _pair(xml)
    prm = ts.harmonic_oscillator_parameters()
    P = 1500
    X = np.zeros((P, 1, 3))
    lam = spec.lambda_vector(ts.harmonic_parameter_sequence(2)[0])
    out = hs.smc_propagate(X, None, lam, prm["timestep"], prm["collision_rate"], kT, 400, seed=11, step0=0, flags=3)
    u = o.reduced_potential(out["x"], ts.harmonic_parameter_sequence(2)[0])
    assert abs(u.mean() - 1.5) < 6 * np.sqrt(1.5 / P) + 0.02          # dt = period / 20 biases <u> by ~1 %
    assert abs(u.std() - np.sqrt(1.5)) < 0.1
    assert np.all(np.isfinite(out["shadow"])) and abs(out["shadow"].mean()) < 0.2
    assert not out["nan"].any()
