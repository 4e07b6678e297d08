"""Parity of the CUDA kernels (through the C ABI) with the oracle and with the reference's golden numbers."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from coddiwomple_b200 import _lib, kernels, testsystems as ts  # noqa: E402
from coddiwomple_b200.system import spec_from_xml  # noqa: E402
from oracle import langevin, resampling, rng, weights  # noqa: E402
from oracle.potential import XmlSystemOracle, kT_kj_per_mol  # noqa: E402
from tests.helpers import (VACUUM_SYSTEMS, cache_frames, eq0000_frames, generic_layer, hybrid_positions, jittered, reference_numbers,  # noqa: E402
                           relative_lambdas, system_xml)

kT = kT_kj_per_mol()


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.fail("the -m gpu tests need a CUDA device: there is no CPU fallback to test")
    from tests import gpu_util
    return gpu_util


def _system(xml):
    from coddiwomple_b200.engine import DeviceSystem
    return XmlSystemOracle(xml), DeviceSystem(xml)


# ------------------------------------------------------------------------------------------------ K1
def test_device_is_blackwell(gpu):
    import ctypes as C
    sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
    _lib.check(_lib.load().cdw_device_info(C.byref(sm), C.byref(ma), C.byref(mi)))
    assert ma.value == 10 and sm.value >= 100


def test_K1_golden_numbers_of_the_reference(gpu):
    """G1-G4 (troubleshooting/openmm_smc.ipynb JSON 260/264/268/339) straight from the GPU kernel."""
    o, ds = _system(system_xml("fluorobenzene"))
    x = eq0000_frames()[:3]
    ref = reference_numbers()
    u0 = gpu.reduced_potential(ds, x, {}, 1 / kT)
    for key, frame in (("frame2", 2), ("frame0", 0), ("frame1", 1)):
        assert u0[frame] == pytest.approx(ref["u_lambda0"][key][0], rel=5e-6)
    dw = gpu.reduced_potential(ds, x, relative_lambdas(1 / 9), 1 / kT) - u0
    _, stats, _ = kernels.weight_stats(dw, always=True)
    assert float(stats[_lib.STAT_MEAN]) == pytest.approx(ref["first_mean_cumulative_work"][0], rel=5e-6)


@pytest.mark.parametrize("name", VACUUM_SYSTEMS)
def test_K1_and_forces_match_oracle(gpu, name):
    o, ds = _system(system_xml(name))
    X = jittered(hybrid_positions(name), 300, seed=1)
    r = np.random.RandomState(2)
    for trial in range(3):
        lam = {k: float(r.uniform(0, 1)) for k in o.global_defaults if k.startswith("lambda")}
        if trial == 0:
            lam = {k: 0.0 for k in lam}
        U, F = o.energy_forces(X, lam)
        u = gpu.reduced_potential(ds, X, lam, 1 / kT)
        np.testing.assert_allclose(u, U / kT, rtol=1e-10)   # north star bar: 1e-6 relative
        e, f = gpu.energy_forces(ds, X, lam)
        np.testing.assert_allclose(e, U, rtol=1e-10)
        assert np.abs(f - F).max() <= 1e-6 * np.abs(F).max()


def test_K1_cache_frames_times_reference_protocol(gpu):
    o, ds = _system(system_xml("fluorobenzene"))
    X = cache_frames().astype(np.float64)
    for q in np.linspace(0, 1, 10):
        lam = relative_lambdas(float(q))
        np.testing.assert_allclose(gpu.reduced_potential(ds, X, lam, 1 / kT), o.reduced_potential(X, lam), rtol=1e-10)


def test_K1_synthetic_systems_and_ragged_sizes(gpu):
    xml, x = ts.alanine_dipeptide_shaped_xml()
    o, ds = _system(xml)
    for P in (1, 3, 5, 127, 1001):   # not multiples of the warps per CTA / the grid
        X = ts.alanine_dipeptide_shaped_ensemble(P).astype(np.float64)
        lam = ts.relative_alchemical_sequence(5)[3]
        np.testing.assert_allclose(gpu.reduced_potential(ds, X, lam, 1 / kT), o.reduced_potential(X, lam), rtol=1e-10)
    oh, dh = _system(ts.harmonic_oscillator_xml())
    X = np.random.RandomState(0).normal(0, 0.1, (1000, 1, 3)).astype(np.float32).astype(np.float64)
    lam = ts.harmonic_parameter_sequence(5)[2]
    np.testing.assert_allclose(gpu.reduced_potential(dh, X, lam, 1 / kT), oh.reduced_potential(X, lam), rtol=1e-12)
    # +1 kJ/mol in U0 raises u by exactly beta (coddiwomple/tests/test_openmm.py:281)
    lam2 = dict(lam)
    lam2[f"{ts.HO_PREFIX}_U0"] += 1.0
    np.testing.assert_allclose(gpu.reduced_potential(dh, X, lam2, 1 / kT) - gpu.reduced_potential(dh, X, lam, 1 / kT), 1 / kT, rtol=1e-9)


def test_K1_empty_batch(gpu):
    _, ds = _system(system_xml("fluorobenzene"))
    assert gpu.reduced_potential(ds, np.zeros((0, 13, 3)), {}, 1.0).shape == (0,)


# ------------------------------------------------------------------------------------------------ K4
@pytest.mark.parametrize("name,n_steps,flags", [("fluorobenzene", 1, 1), ("fluorobenzene", 3, 1), ("fluorobenzene", 2, 3), ("ethylbenzene", 1, 1),
                                                ("ethylbenzene", 2, 3), ("aniline", 2, 0), ("ala", 1, 1), ("ala", 2, 0), ("ala", 2, 3)])
def test_K4_baoab_matches_oracle(gpu, name, n_steps, flags):
    """"ala" is the 22-atom AlanineDipeptideVacuum-shaped system the headline benchmark is quoted on (team of 4 lanes)."""
    o, ds = _system(ts.alanine_dipeptide_shaped_xml()[0] if name == "ala" else system_xml(name))
    P = 96
    if name == "fluorobenzene":
        X = cache_frames()[:P].astype(np.float64)
    elif name == "ala":
        X = ts.alanine_dipeptide_shaped_ensemble(P).astype(np.float64)
    else:
        X = np.repeat(hybrid_positions(name)[None], P, 0).astype(np.float32).astype(np.float64)
    lam = relative_lambdas(0.3)
    V = np.random.RandomState(1).normal(0, 0.3, X.shape)
    langevin._rattle_v(X, V, 1.0 / o.masses, o.constraints, 1e-6)
    V = V.astype(np.float32).astype(np.float64)
    ref = langevin.baoab(lambda xx: o.energy_forces(xx, lam), X, V, o.masses, o.constraints, 0.002, 1.0, kT, n_steps, seed=7, gids=np.arange(5, 5 + P),
                         step0=11, randomize_velocities=bool(flags & 1), track_works=bool(flags & 2))
    out = gpu.smc_propagate(ds, X, V, lam, 0.002, 1.0, kT, n_steps, seed=7, step0=11, gid0=5, flags=flags, aux=np.full(P, 0.25), cum=np.full(P, 1.5))
    assert np.abs(out["x"] - ref["x"]).max() < 2e-7
    assert np.abs(out["v"] - ref["v"]).max() < 2e-5
    np.testing.assert_allclose(out["u_first"], ref["u_first"] / kT, rtol=1e-10)
    np.testing.assert_allclose(out["u_last"], ref["u_last"] / kT, atol=2e-4)
    np.testing.assert_allclose(out["inc"], ref["u_first"] / kT + 0.25, rtol=1e-10)
    np.testing.assert_allclose(out["w"], 1.5 + ref["u_first"] / kT + 0.25, rtol=1e-10)
    np.testing.assert_array_equal(out["aux"], -out["u_last"])
    if flags & 2:
        np.testing.assert_allclose(out["shadow"], ref["shadow_work"] / kT, atol=3e-4)
        np.testing.assert_allclose(out["proposal"], ref["proposal_work"] / kT, atol=2e-5)
    for (i, j, d) in o.constraints:
        assert np.abs(np.linalg.norm(out["x"][:, i] - out["x"][:, j], axis=-1) / d - 1).max() < 3e-6
    # the running minimum of W that feeds the log-sum-exp
    key = out["wmin"][0]
    wmin = out["w"].min()
    b = np.float64(wmin).view(np.uint64)
    assert key == (~b if b >> np.uint64(63) else b | np.uint64(1 << 63))


@pytest.mark.parametrize("team", [1, 2, 8])
def test_K4_other_team_sizes_match_oracle(gpu, team):
    """cdw_system_set_team: teams of 2 (what the engines pin for ensembles that fill the GPU several times over), 1 and 8 lanes per
    replica against the oracle, and the look-ahead launch against the plain one at that team size."""
    o, ds = _system(system_xml("fluorobenzene"))
    _lib.check(ds.lib.cdw_system_set_team(ds.handle, team))
    assert ds.lib.cdw_system_team(ds.handle) == team
    P = 70
    X = cache_frames()[:P].astype(np.float64)
    lam = relative_lambdas(0.3)
    V = np.zeros_like(X)
    ref = langevin.baoab(lambda xx: o.energy_forces(xx, lam), X, V, o.masses, o.constraints, 0.002, 1.0, kT, 2, seed=7, gids=np.arange(5, 5 + P),
                         step0=11, randomize_velocities=True)
    out = gpu.smc_propagate(ds, X, V, lam, 0.002, 1.0, kT, 2, seed=7, step0=11, gid0=5, flags=1, aux=np.full(P, 0.25), cum=np.full(P, 1.5))
    assert np.abs(out["x"] - ref["x"]).max() < 2e-7 and np.abs(out["v"] - ref["v"]).max() < 2e-5
    np.testing.assert_allclose(out["u_first"], ref["u_first"] / kT, rtol=1e-10)
    np.testing.assert_allclose(out["inc"], ref["u_first"] / kT + 0.25, rtol=1e-10)
    e, f = gpu.energy_forces(ds, X, lam)
    U, F = o.energy_forces(X, lam)
    np.testing.assert_allclose(e, U, rtol=1e-10)
    assert np.abs(f - F).max() <= 1e-6 * np.abs(F).max()


@pytest.mark.parametrize("splitting,dt", [("O V R V O", 0.002), ("O { V R V } O", 0.004), ("{ V R O R V }", 0.006)])
def test_K4_splitting_programs_match_oracle(gpu, splitting, dt):
    """cdw_langevin_params.splitting: OMMLI's splitting strings incl. the Metropolised block (openmm/integrators.py:207-217, 427-444)
    against the oracle interpreter: same trials, same acceptances, same trajectories and works."""
    o, ds = _system(system_xml("fluorobenzene"))
    P = 64
    lam = relative_lambdas(0.3)
    X = gpu.smc_propagate(ds, cache_frames()[:P].astype(np.float64), None, lam, 0.002, 1.0, kT, 2, seed=1, step0=1, flags=1)["x"]   # onto the constraint manifold
    V = np.zeros_like(X)
    ref = langevin.langevin_program(lambda xx: o.energy_forces(xx, lam), X, V, o.masses, o.constraints, dt, 5.0, kT, 3, seed=7, gids=np.arange(5, 5 + P),
                                    step0=11, splitting=splitting, randomize_velocities=True)
    out = gpu.smc_propagate(ds, X, V, lam, dt, 5.0, kT, 3, seed=7, step0=11, gid0=5, flags=1, aux=np.zeros(P), cum=np.zeros(P),
                            splitting=langevin.encode_splitting(splitting))
    np.testing.assert_array_equal(out["trials"][:, 0], ref["trials"])
    np.testing.assert_array_equal(out["trials"][:, 1], ref["accepted"])
    assert np.abs(out["x"] - ref["x"]).max() < 5e-7 and np.abs(out["v"] - ref["v"]).max() < 5e-5
    np.testing.assert_allclose(out["u_last"], ref["u_last"] / kT, atol=5e-4)
    np.testing.assert_allclose(out["shadow"], ref["shadow_work"] / kT, atol=1e-3)
    np.testing.assert_allclose(out["proposal"], ref["proposal_work"] / kT, atol=1e-4)
    if "{" in splitting:
        assert 0 < out["trials"][:, 1].sum() <= out["trials"][:, 0].sum() == 3 * P


def test_lookahead_launch_matches_the_plain_launch_and_reverts_bad_moves(gpu):
    """cdw_smc_lookahead against cdw_smc_propagate (which is checked against the oracle above): same states bit for bit, the next
    incremental works equal the opening works of a plain launch at the next lambda; a replica whose move ends in a non-finite energy
    (here: an infinite input velocity) is reverted in the middle of a warp of healthy ones (openmm/propagators.py:166-186)."""
    o, ds = _system(system_xml("fluorobenzene"))
    P = 100
    X = cache_frames()[:P].astype(np.float64)
    V = np.zeros_like(X)
    lam1, lam2 = relative_lambdas(0.3), relative_lambdas(0.4)
    eq = gpu.smc_lookahead(ds, X, V, None, lam1, lam1, 0.002, 1.0, kT, 0, seed=7, step0=0)
    np.testing.assert_allclose(eq["aux"], -o.reduced_potential(X, lam1), rtol=1e-10)
    a = gpu.smc_lookahead(ds, X, V, eq["force"], lam1, lam2, 0.002, 1.0, kT, 2, seed=7, step0=2, flags=1, labels=np.arange(P) + 3)
    b = gpu.smc_propagate(ds, X, V, lam1, 0.002, 1.0, kT, 2, seed=7, step0=2, flags=1, aux=np.zeros(P), cum=np.zeros(P))
    assert np.abs(a["x"] - b["x"]).max() < 5e-8 and np.abs(a["v"] - b["v"]).max() < 5e-6   # (static / dynamic split vs one pass: summation order)
    np.testing.assert_allclose(a["aux"], b["aux"], atol=2e-4)               # (the energies follow the fp32-ulp position differences ...)
    np.testing.assert_allclose(a["aux"], -o.reduced_potential(a["x"], lam1), rtol=1e-10)   # (... and are exact at the launch's own rows)
    nxt = gpu.smc_propagate(ds, a["x"], a["v"], lam2, 0.002, 1.0, kT, 0, seed=7, step0=4, aux=a["aux"], cum=np.zeros(P))
    np.testing.assert_allclose(a["inc_next"], nxt["inc"], rtol=1e-9, atol=1e-11)
    np.testing.assert_array_equal(a["label"], np.arange(P) + 3)
    assert not a["nan"].any()
    Vb = V.copy()
    Vb[[5, 37, 38], 0, 0] = np.inf
    c = gpu.smc_lookahead(ds, X, Vb, eq["force"], lam1, lam2, 0.002, 1.0, kT, 2, seed=7, step0=2)
    assert sorted(np.nonzero(c["nan"])[0]) == [5, 37, 38]
    np.testing.assert_array_equal(c["x"][[5, 37, 38]], X[[5, 37, 38]])
    keep = gpu.smc_lookahead(ds, X, V, None, lam1, lam2, 0.002, 1.0, kT, 0, seed=7, step0=0)
    np.testing.assert_allclose(c["aux"][[5, 37, 38]], keep["aux"][[5, 37, 38]], rtol=1e-13)
    np.testing.assert_allclose(c["inc_next"][[5, 37, 38]], keep["inc_next"][[5, 37, 38]], rtol=1e-9, atol=1e-11)
    good = np.setdiff1d(np.arange(P), [5, 37, 38])
    d = gpu.smc_lookahead(ds, X, V, eq["force"], lam1, lam2, 0.002, 1.0, kT, 2, seed=7, step0=2)
    np.testing.assert_array_equal(c["pos"][good], d["pos"][good])         # the neighbours of a reverted replica are untouched (bitwise)
    np.testing.assert_array_equal(c["inc_next"][good], d["inc_next"][good])


def test_K4_gather_on_load_nan_guard_and_determinism(gpu):
    o, ds = _system(system_xml("fluorobenzene"))
    X = cache_frames()[:64].astype(np.float64)
    anc = np.random.RandomState(0).randint(0, 64, 64)
    aux = np.arange(64) * 0.5
    labels = np.arange(64) + 100
    lam = relative_lambdas(0.5)
    out = gpu.smc_propagate(ds, X, np.zeros_like(X), lam, 0.002, 1.0, kT, 0, seed=1, step0=0, anc=anc, aux=aux, cum=np.zeros(64), labels=labels)
    np.testing.assert_array_equal(out["x"], X[anc])
    np.testing.assert_array_equal(out["label"], labels[anc])
    u = o.reduced_potential(X[anc], lam)
    np.testing.assert_allclose(out["inc"], u + aux[anc], rtol=1e-10)
    # NaN guard
    Xb = X.copy()
    Xb[7, 12] = Xb[7, 5]
    a = gpu.smc_propagate(ds, Xb, np.zeros_like(X), lam, 0.002, 1.0, kT, 1, seed=1, step0=1, aux=np.zeros(64), cum=np.zeros(64))
    assert a["nan"].sum() == 1 and a["nan"][7] == 1
    np.testing.assert_array_equal(a["x"][7], Xb[7])
    # bitwise run-to-run determinism (counter-based noise, no atomics on floats)
    b = gpu.smc_propagate(ds, Xb, np.zeros_like(X), lam, 0.002, 1.0, kT, 1, seed=1, step0=1, aux=np.zeros(64), cum=np.zeros(64))
    for k in ("pos", "vel", "inc", "aux"):
        np.testing.assert_array_equal(a[k], b[k])
    # noise is keyed by the GLOBAL particle id: a shard starting at gid0 = 32 reproduces particles 32.. of the full run
    full = gpu.smc_propagate(ds, X, None, lam, 0.002, 1.0, kT, 2, seed=9, step0=4, gid0=0, flags=1)
    shard = gpu.smc_propagate(ds, X[32:], None, lam, 0.002, 1.0, kT, 2, seed=9, step0=4, gid0=32, flags=1)
    np.testing.assert_array_equal(full["pos"][32:], shard["pos"])
    np.testing.assert_array_equal(full["vel"][32:], shard["vel"])


# ------------------------------------------------------------------------------------------------ K2 / K3
@pytest.mark.parametrize("case", range(6))
def test_K2_K3_reproduce_the_reference_generic_layer(gpu, case):
    """nESS, the normaliser, np.random.choice's ancestors and the post-resampling works recorded from the reference's
    own Resampler / nESS (tests/golden/generic_layer.json), reproduced by the kernels."""
    c = generic_layer()["cases"][case]
    cum, inc = np.array(c["cumulative_works_in"]), np.array(c["incremental_works"])
    W, stats, ws = kernels.weight_stats(cum, incremental=inc, floor_threshold=0.5)
    s = stats.cpu().numpy()
    assert s[_lib.STAT_NESS] == pytest.approx(c["nESS"], rel=1e-12)
    assert bool(s[_lib.STAT_RESAMPLE]) == c["threshold_says_resample"]
    np.testing.assert_array_equal(W.cpu().numpy(), cum + inc)
    anc = kernels.resample_indices(cum + inc, np.array(c["uniforms"]), kind="multinomial")
    np.testing.assert_array_equal(anc, np.array(c["after_forced"]["state_label"]))   # == np.random.choice with the same seed
    mean = float(kernels.weight_stats(cum + inc, always=True)[1][_lib.STAT_MEAN])
    np.testing.assert_allclose(cum + (mean - cum), np.array(c["after_forced"]["cumulative_work"]), rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("n,scale", [(1, 1.0), (2, 1.0), (1000, 0.1), (1000, 3.0), (2047, 1.0), (2048, 1.0), (2049, 1.0), (10 ** 5, 1.0), (10 ** 6, 0.1),
                                     (10 ** 6, 3.0)])
def test_K3_ancestors_bit_exact_vs_oracle(gpu, n, scale):
    r = np.random.RandomState(n % 1000 + int(scale * 10))
    W = r.normal(0, scale, n)
    u = r.random_sample(n)
    np.testing.assert_array_equal(kernels.resample_indices(W, u, kind="multinomial"), resampling.ancestors_fixed(W, u))
    np.testing.assert_array_equal(kernels.resample_indices(W, np.array([0.4321]), kind="systematic"),
                                  resampling.ancestors(W, None, kind="systematic", u0=0.4321))
    if n <= 10 ** 5:   # and with numpy's float path, i.e. what the reference's np.random.choice returns
        np.testing.assert_array_equal(kernels.resample_indices(W, u, kind="multinomial"), resampling.multinomial_float(W, u))


def test_K3_extreme_weights(gpu):
    W = np.full(5000, 800.0)   # every weight but one underflows: all ancestors must be the survivor
    W[1234] = 0.0
    u = np.random.RandomState(0).random_sample(5000)
    assert np.all(kernels.resample_indices(W, u) == 1234)
    W = np.zeros(4096)         # exactly flat weights: total = 2^61, ancestors = floor(u N)
    np.testing.assert_array_equal(kernels.resample_indices(W, u[:4096]), np.floor(u[:4096] * 4096).astype(np.int64))
    np.testing.assert_array_equal(kernels.resample_indices(W, np.array([0.5]), kind="systematic"), np.arange(4096))


@pytest.mark.parametrize("n", [10 ** 7])
def test_K3_large_properties(gpu, n):
    """Size-independent properties at sweep sizes: systematic ancestors are sorted and |N_i - N w_i| < 1; the multinomial
    ancestors equal the oracle's (numpy, ~10 s); philox uniforms equal the oracle's stream."""
    W = torch.randn(n, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    a = kernels.resample_indices(W, torch.tensor([0.7], dtype=torch.float64), kind="systematic")
    assert bool((a[1:] >= a[:-1]).all())
    counts = torch.bincount(a.long(), minlength=n).double()
    w = torch.softmax(-W, 0)
    assert float((counts - n * w).abs().max()) < 1.0 + 1e-6
    u = kernels.philox_uniforms(6, 1, n)
    np.testing.assert_array_equal(u[:100000].cpu().numpy(), rng.uniforms53(6, n, stream=1)[:100000])
    am = kernels.resample_indices(W, u, kind="multinomial").cpu().numpy()
    np.testing.assert_array_equal(am, resampling.ancestors_fixed(W.cpu().numpy(), u.cpu().numpy()))
    # the slot ranges the scan emits (floating-point estimate, exact forward map only near an integer) against the exact integer search
    for u0 in (0.7, 0.0, 1.0 - 2.0 ** -53):
        a = kernels.resample_indices(W, torch.tensor([u0], dtype=torch.float64), kind="systematic").cpu().numpy()
        np.testing.assert_array_equal(a, resampling.ancestors(W.cpu().numpy(), None, kind="systematic", u0=u0))


def test_K3b_gather_bit_exact(gpu):
    r = np.random.RandomState(0)
    for P, A in ((1, 13), (1000, 13), (4097, 22), (50000, 1)):
        pos = torch.as_tensor(r.normal(0, 1, (P, A, 4)).astype(np.float32)).cuda()
        vel = torch.as_tensor(r.normal(0, 1, (P, A, 4)).astype(np.float32)).cuda()
        aux = torch.as_tensor(r.normal(0, 1, P)).cuda()
        lab = torch.arange(P, dtype=torch.int32).cuda()
        anc = torch.as_tensor(r.randint(0, P, P).astype(np.int32)).cuda()
        po, vo, ao, lo = kernels.gather_particles(pos, vel, anc, aux, lab)
        idx = anc.long()
        assert torch.equal(po, pos[idx]) and torch.equal(vo, vel[idx]) and torch.equal(ao, aux[idx]) and torch.equal(lo, lab[idx])


def test_K3b_peer_gather_tma_bit_exact(gpu):
    """cdw_gather_particles_peer (bulk row copies through shared memory, one bulk store per chunk) with the local buffers standing in for two
    "ranks": global ancestor ids resolve to (rank, local row); bit-exact, ragged last chunk, with and without aux / labels."""
    import ctypes as C
    lib = _lib.load()
    r = np.random.RandomState(1)
    for n_per, A in ((33, 13), (1000, 13), (4097, 22), (257, 1)):
        bufs = [dict(pos=torch.as_tensor(r.normal(0, 1, (n_per, A, 4)).astype(np.float32)).cuda(), vel=torch.as_tensor(r.normal(0, 1, (n_per, A, 4)).astype(np.float32)).cuda(),
                     aux=torch.as_tensor(r.normal(0, 1, n_per)).cuda(), lab=torch.as_tensor(r.randint(0, 10 ** 6, n_per).astype(np.int32)).cuda()) for _ in range(2)]
        ptrs = {k: torch.tensor([b[k].data_ptr() for b in bufs], dtype=torch.int64, device="cuda") for k in ("pos", "vel", "aux", "lab")}
        n_local = n_per + 7
        anc = torch.as_tensor(r.randint(0, 2 * n_per, n_local).astype(np.int32)).cuda()
        cat = {k: torch.cat([b[k] for b in bufs]) for k in ("pos", "vel", "aux", "lab")}
        for with_meta in (True, False):
            po, vo = torch.zeros(n_local, A, 4, device="cuda"), torch.zeros(n_local, A, 4, device="cuda")
            ao, lo = torch.zeros(n_local, dtype=torch.float64, device="cuda"), torch.zeros(n_local, dtype=torch.int32, device="cuda")
            _lib.check(lib.cdw_gather_particles_peer(_lib.ptr(ptrs["pos"]), _lib.ptr(ptrs["vel"]), _lib.ptr(ptrs["aux"]) if with_meta else None,
                                                     _lib.ptr(ptrs["lab"]) if with_meta else None, 2, n_per, _lib.ptr(anc), _lib.ptr(po), _lib.ptr(vo),
                                                     _lib.ptr(ao) if with_meta else None, _lib.ptr(lo) if with_meta else None, n_local, A, _lib.current_stream()),
                       "cdw_gather_particles_peer")
            idx = anc.long()
            assert torch.equal(po, cat["pos"][idx]) and torch.equal(vo, cat["vel"][idx])
            if with_meta:
                assert torch.equal(ao, cat["aux"][idx]) and torch.equal(lo, cat["lab"][idx])


def test_K2_large_reduction_accuracy(gpu):
    n = 3_000_001
    W = np.random.RandomState(1).normal(0, 2.0, n)
    _, stats, _ = kernels.weight_stats(W, floor_threshold=0.5)
    s = stats.cpu().numpy()
    assert s[_lib.STAT_NESS] == pytest.approx(weights.nESS(W), rel=1e-11)
    from scipy.special import logsumexp
    assert s[_lib.STAT_MEAN] == pytest.approx(-logsumexp(-W) + np.log(n), rel=1e-12, abs=1e-12)
    assert s[_lib.STAT_WMIN] == W.min()
    s2 = kernels.weight_stats(W, floor_threshold=0.5)[1].cpu().numpy()
    np.testing.assert_array_equal(s, s2)   # run-to-run deterministic reduction


@pytest.mark.parametrize("n,kind", [(1, 0), (7, 1), (1000, 0), (10000, 0), (10000, 1), (24576, 0)])
def test_resample_statistics_are_exact_integer_sums(gpu, n, kind):
    """cdw_smc_resample: sum e and sum e^2 are 128-bit integer sums of the quantised exponentials (oracle.resampling.
    canonical_statistics): bit for bit the oracle's, whatever the grid; nESS follows bitwise, the normaliser to 1e-13 (the
    device's own log); the ancestors equal the oracle's; the workspace is left at rest."""
    import ctypes as C
    from coddiwomple_b200 import kernels
    lib = _lib.load()
    W_host = np.random.RandomState(n).normal(0.0, 1.5, n)
    W = torch.as_tensor(W_host).cuda()
    ws = kernels.WeightWorkspace(n)
    ws.wmin_key.fill_(-1)
    Wc = torch.empty_like(W)
    _lib.check(lib.cdw_weight_update(_lib.ptr(W), None, None, None, n, None, _lib.ptr(Wc), _lib.ptr(ws.wmin_key), _lib.current_stream()))
    cum = torch.zeros_like(W)
    anc = torch.zeros(n, dtype=torch.int32, device="cuda")
    u = torch.zeros(n, dtype=torch.float64, device="cuda")
    _lib.check(lib.cdw_smc_resample(kind, _lib.ptr(W), n, -1.0, _lib.ptr(ws.wmin_key), C.c_uint64(11), C.c_uint64(3), _lib.ptr(ws.stats), _lib.ptr(cum),
                                    _lib.ptr(cum), _lib.ptr(anc), _lib.ptr(ws.cdf), _lib.ptr(u), _lib.ptr(ws.buf), ws.bytes, _lib.current_stream()))
    s = ws.stats.cpu().numpy()
    wmin, sum_e, sum_e2, ness, mean = resampling.canonical_statistics(W_host)
    assert s[_lib.STAT_WMIN] == wmin and s[_lib.STAT_SUM_E] == sum_e and s[_lib.STAT_SUM_E2] == sum_e2 and s[_lib.STAT_NESS] == ness
    assert s[_lib.STAT_MEAN] == pytest.approx(mean, rel=1e-13, abs=1e-13)
    uni = u.cpu().numpy()
    ref = resampling.ancestors(W_host, uni if kind == 0 else None, kind="multinomial" if kind == 0 else "systematic", u0=None if kind == 0 else uni[0])
    np.testing.assert_array_equal(anc.cpu().numpy(), ref)
    np.testing.assert_allclose(cum.cpu().numpy(), np.full(n, s[_lib.STAT_MEAN]), rtol=1e-15, atol=1e-15)
    assert int(ws.wmin_key.item()) == -1                     # re-armed
    hdr = ws.buf[:8].cpu().numpy()
    assert not hdr[[0, 1, 2, 3, 4, 7]].any()                 # accumulators and tickets back at rest
    q, k = resampling.fixed_point_weights(W_host)
    total = int(q.sum(dtype=object))
    assert int(hdr[5]) == k and int(hdr[6]) == total and s[_lib.STAT_TOTAL] == float(total)


def test_weigh_resample_composes_the_weights_through_the_previous_ancestors(gpu):
    """cdw_smc_weigh_resample (the resampling of the look-ahead loop): W[i] = cum[i] + inc_src[anc_prev[i]], then exactly
    cdw_smc_resample on W."""
    import ctypes as C
    from coddiwomple_b200 import kernels
    lib = _lib.load()
    n = 5000
    r = np.random.RandomState(4)
    inc_src, cum0, prev = r.normal(0, 1.0, n), r.normal(0, 0.5, n), r.randint(0, n, n).astype(np.int32)
    for anc_prev in (prev, None):
        ws = kernels.WeightWorkspace(n)
        ws.wmin_key.fill_(-1)
        d = lambda a, dt=torch.float64: torch.as_tensor(a, dtype=dt).cuda()
        cum, inc, W, u = d(cum0), torch.zeros(n, dtype=torch.float64, device="cuda"), torch.zeros(n, dtype=torch.float64, device="cuda"), torch.zeros(n, dtype=torch.float64, device="cuda")
        anc = torch.zeros(n, dtype=torch.int32, device="cuda")
        ap = None if anc_prev is None else d(anc_prev, torch.int32)
        _lib.check(lib.cdw_smc_weigh_resample(0, _lib.ptr(d(inc_src)), _lib.ptr(ap), n, 0.5, _lib.ptr(ws.wmin_key), C.c_uint64(5), C.c_uint64(2), _lib.ptr(ws.stats),
                                              _lib.ptr(cum), _lib.ptr(inc), _lib.ptr(W), _lib.ptr(anc), _lib.ptr(ws.cdf), _lib.ptr(u), _lib.ptr(ws.buf), ws.bytes,
                                              _lib.current_stream()))
        inc_ref = inc_src if anc_prev is None else inc_src[anc_prev]
        np.testing.assert_array_equal(inc.cpu().numpy(), inc_ref)
        np.testing.assert_array_equal(W.cpu().numpy(), cum0 + inc_ref)
        ref = weights.resample_step(cum0, inc_ref, np.zeros(n), np.arange(n), uniforms=u.cpu().numpy(), kind="multinomial", floor_threshold=0.5)
        np.testing.assert_array_equal(anc.cpu().numpy(), ref["ancestors"])
        np.testing.assert_allclose(cum.cpu().numpy(), ref["cum"], rtol=1e-12, atol=1e-12)
        assert bool(ws.stats[_lib.STAT_RESAMPLE].item()) == ref["resampled"]


@pytest.mark.parametrize("n,kind", [(50000, 0), (50000, 1), (100003, 0)])
def test_multi_kernel_resample_draws_its_own_uniforms(gpu, n, kind):
    """cdw_smc_resample at sizes of several tiles (statistics + tile sums + tile scan with emission + the lookup that draws the
    Philox uniforms itself): the uniforms are the documented counter-based stream and the ancestors are the oracle's."""
    import ctypes as C
    from coddiwomple_b200 import kernels
    lib = _lib.load()
    W_host = np.random.RandomState(n + kind).normal(0.0, 1.5, n)
    W = torch.as_tensor(W_host).cuda()
    ws = kernels.WeightWorkspace(n)
    ws.wmin_key.fill_(-1)
    Wc = torch.empty_like(W)
    _lib.check(lib.cdw_weight_update(_lib.ptr(W), None, None, None, n, None, _lib.ptr(Wc), _lib.ptr(ws.wmin_key), _lib.current_stream()))
    cum = torch.zeros_like(W)
    anc = torch.zeros(n, dtype=torch.int32, device="cuda")
    u = torch.zeros(n, dtype=torch.float64, device="cuda")
    for rep in range(2):   # twice on the same workspace: the kernels leave it at rest
        _lib.check(lib.cdw_smc_resample(kind, _lib.ptr(W), n, -1.0, _lib.ptr(ws.wmin_key), C.c_uint64(11), C.c_uint64(3 + rep), _lib.ptr(ws.stats),
                                        _lib.ptr(cum), _lib.ptr(cum), _lib.ptr(anc), _lib.ptr(ws.cdf), _lib.ptr(u), _lib.ptr(ws.buf), ws.bytes,
                                        _lib.current_stream()))
        assert int(ws.wmin_key.item()) == -1
        _lib.check(lib.cdw_weight_update(_lib.ptr(W), None, None, None, n, None, _lib.ptr(Wc), _lib.ptr(ws.wmin_key), _lib.current_stream()))
        uni = u.cpu().numpy()
        n_u = n if kind == 0 else 1
        np.testing.assert_array_equal(uni[:n_u], rng.uniforms53(11, n_u, stream=3 + rep))
        ref = resampling.ancestors(W_host, uni if kind == 0 else None, kind="multinomial" if kind == 0 else "systematic", u0=None if kind == 0 else uni[0])
        np.testing.assert_array_equal(anc.cpu().numpy(), ref)
    s = ws.stats.cpu().numpy()
    assert s[_lib.STAT_NESS] == pytest.approx(weights.nESS(W_host), rel=1e-12)
