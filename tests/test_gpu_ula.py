"""Euler-Maruyama (ULA) proposal on the GPU (CDW_FLAG_ULA): kernel vs oracle, the engine vs the oracle anneal, and the
free-energy anchors (harmonic oscillator Delta F = 1 kT; the 3-d analogue of the reference's toy G8:
troubleshooting/test_csmc.py:176-226, F = -log sqrt 2 per dimension)."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from coddiwomple_b200 import _lib, testsystems as ts  # noqa: E402
from oracle import smc as oracle_smc, ula  # noqa: E402
from oracle.potential import XmlSystemOracle, kT_kj_per_mol  # noqa: E402
from tests.helpers import cache_frames, hybrid_positions, relative_lambdas, system_xml  # noqa: E402

kT = kT_kj_per_mol()


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("the -m gpu tests need a CUDA device: there is no CPU fallback to test")


def _propagate(xml, X, lam, dt, n_steps, seed, step0, gid0, aux, cum):
    from coddiwomple_b200.engine import DeviceSystem, LangevinParameters, positions_to_rows
    ds = DeviceSystem(xml)
    lib = ds.lib
    P = X.shape[0]
    pos = positions_to_rows(X)
    pos_o = torch.empty_like(pos)
    dev = lambda a: torch.as_tensor(a, dtype=torch.float64).cuda()
    o = {k: torch.zeros(P, dtype=torch.float64, device="cuda") for k in ("inc", "w", "aux", "u_first", "u_last", "proposal")}
    nan = torch.zeros(P, dtype=torch.int32, device="cuda")
    key = torch.full((1,), -1, dtype=torch.int64, device="cuda")
    lamv = torch.as_tensor(ds.spec.lambda_vector(lam)).cuda()
    prm = LangevinParameters(timestep=dt, collision_rate=0.0, n_steps=n_steps).as_struct(_lib.CDW_FLAG_ULA)
    aux_d, cum_d = dev(aux), dev(cum)
    _lib.check(lib.cdw_smc_propagate(ds.handle, _lib.ptr(pos), None, _lib.ptr(pos_o), None, P, None, _lib.ptr(aux_d), _lib.ptr(cum_d), None, _lib.ptr(lamv),
                                     C.byref(prm), C.c_uint64(seed), C.c_uint64(step0), C.c_int64(gid0), _lib.ptr(o["inc"]), _lib.ptr(o["w"]), _lib.ptr(o["aux"]),
                                     None, _lib.ptr(o["u_first"]), _lib.ptr(o["u_last"]), None, _lib.ptr(o["proposal"]), _lib.ptr(nan), _lib.ptr(key),
                                     _lib.current_stream()), "cdw_smc_propagate")
    out = {k: v.cpu().numpy() for k, v in o.items()}
    out["x"] = pos_o[..., :3].double().cpu().numpy()
    out["nan"] = nan.cpu().numpy()
    return out


@pytest.mark.parametrize("name,n_steps,P", [("fluorobenzene", 1, 100), ("fluorobenzene", 3, 37), ("ethylbenzene", 2, 64), ("ala", 1, 64), ("ala", 4, 33)])
def test_kernel_matches_oracle(name, n_steps, P):
    xml = ts.alanine_dipeptide_shaped_xml()[0] if name == "ala" else system_xml(name)
    o = XmlSystemOracle(xml)
    if name == "fluorobenzene":
        X = cache_frames()[np.random.RandomState(0).randint(0, 151, P)].astype(np.float64)
    elif name == "ala":
        X = ts.alanine_dipeptide_shaped_ensemble(P).astype(np.float64)
    else:
        X = np.repeat(hybrid_positions(name)[None], P, 0).astype(np.float32).astype(np.float64)
    lam = relative_lambdas(0.3)
    dt = 0.0005
    ref = ula.ula(lambda xx: o.energy_forces(xx, lam), X, o.masses, o.constraints, dt, kT, n_steps, seed=7, gids=np.arange(5, 5 + P), step0=11)
    out = _propagate(xml, X, lam, dt, n_steps, 7, 11, 5, np.full(P, 0.25), np.full(P, 1.5))
    assert np.abs(out["x"] - ref["x"]).max() < 2e-7
    np.testing.assert_allclose(out["u_first"], ref["u_first"] / kT, rtol=1e-12)
    np.testing.assert_allclose(out["u_last"], ref["u_last"] / kT, atol=2e-4)
    np.testing.assert_allclose(out["proposal"], ref["proposal_work"], atol=2e-4)
    np.testing.assert_allclose(out["inc"], out["u_last"] + 0.25 + out["proposal"], rtol=1e-12)
    np.testing.assert_allclose(out["w"], 1.5 + out["inc"], rtol=1e-12)
    np.testing.assert_array_equal(out["aux"], -out["u_last"])
    assert not out["nan"].any()
    for (i, j, d) in o.constraints:
        assert np.abs(np.linalg.norm(out["x"][:, i] - out["x"][:, j], axis=-1) / d - 1).max() < 3e-6


def test_engine_first_iterations_track_the_oracle_anneal():
    from coddiwomple_b200.engine import LangevinParameters, SMCEngine
    frames = cache_frames()
    x0 = frames[np.random.RandomState(1).randint(0, len(frames), 128)]
    seq = ts.relative_alchemical_sequence(6)
    eng = SMCEngine(system_xml("fluorobenzene"), 128, seq, langevin=LangevinParameters(timestep=0.0005), seed=5, resample_seed=6, record_history=True,
                    proposal="ula")
    eng.anneal(x0)
    trace = []
    out = oracle_smc.anneal(XmlSystemOracle(system_xml("fluorobenzene")), x0, seq, dt=0.0005, seed=5, resample_seed=6, trace=trace, proposal="ula")
    h = eng.ensemble.history
    np.testing.assert_allclose(h[0]["inc"].cpu().numpy(), trace[0]["inc"], atol=5e-4)
    np.testing.assert_array_equal(h[0]["anc"].cpu().numpy(), trace[0]["ancestors"])
    assert abs(eng.free_energy() - out["free_energy"]) < 0.1


def test_harmonic_oscillator_free_energies():
    from coddiwomple_b200.engine import LangevinParameters, SMCEngine
    prm = ts.harmonic_oscillator_parameters()
    P = 20000
    x0 = np.random.RandomState(0).normal(0, prm["sigma"], (P, 1, 3)).astype(np.float32)
    # (i) the reference's HO protocol: Delta F = 1 kT exactly (tests/utils.py:64-65)
    eng = SMCEngine(ts.harmonic_oscillator_xml(), P, ts.harmonic_parameter_sequence(41), langevin=LangevinParameters(timestep=0.2), seed=1, resample_seed=2,
                    proposal="ula")
    eng.anneal(x0)
    assert abs(eng.free_energy() - 1.0) < 0.03, eng.free_energy()
    # (ii) variance doubling, the reference's toy G8 in three dimensions: F = -3 log sqrt 2 (test_csmc.py:19-24, 176-226)
    K = prm["K"]
    seq = [{f"{ts.HO_PREFIX}_K": K * (1.0 - 0.5 * q)} for q in np.linspace(0, 1, 10)]
    for always in (False, True):   # SMC and (threshold below every nESS) plain SIS
        eng = SMCEngine(ts.harmonic_oscillator_xml(), P, seq, langevin=LangevinParameters(timestep=0.4), seed=3, resample_seed=4, proposal="ula",
                        floor_threshold=0.5 if not always else 0.0)
        eng.anneal(x0)
        assert abs(eng.free_energy() + 1.5 * np.log(2.0)) < 0.03, eng.free_energy()
        assert int(eng.ensemble.nan_flags.sum()) == 0


def test_OMMEMP_apply_returns_the_kernel_log_ratio():
    from coddiwomple_b200.openmm.integrators import OpenMMEulerMaruyamaIntegrator
    from coddiwomple_b200.openmm.propagators import OMMEMP
    from coddiwomple_b200.openmm.states import OpenMMParticleState, OpenMMPDFState
    xml = system_xml("fluorobenzene")
    o = XmlSystemOracle(xml)
    from coddiwomple_b200.openmm.states import RelativeAlchemicalState
    pdf = OpenMMPDFState(xml, RelativeAlchemicalState, temperature=300.0, pressure=None)
    lam = relative_lambdas(0.4)
    pdf.set_parameters(lam)
    integ = OpenMMEulerMaruyamaIntegrator(timestep=0.0005, temperature=300.0, mass_vector=o.masses)
    prop = OMMEMP(pdf, integ, seed=9)
    x_old = cache_frames()[3].astype(np.float64)
    ps = OpenMMParticleState(positions=x_old.copy())
    ps, work = prop.apply(ps, n_steps=1, apply_pdf_to_context=True)
    x_new = np.asarray(ps.positions, dtype=np.float64)
    tau = integ.tau()[:, None]
    F_old, F_new = o.energy_forces(x_old[None], lam)[1][0], o.energy_forces(x_new[None], lam)[1][0]
    fwd = -(((x_new - x_old - tau * F_old / kT) ** 2) / (4 * tau)).sum()
    bwd = -(((x_old - x_new - tau * F_new / kT) ** 2) / (4 * tau)).sum()
    assert abs(work - (fwd - bwd)) < 5e-4     # integrators.py:864-888 on the move the kernel made
    assert integ.get_proposal_work() == work


# ------------------------------------------------------------------------------------------------ controlled SMC: the quadratic twist
@pytest.mark.parametrize("name,n_steps,P,ref_norm", [("fluorobenzene", 1, 100, False), ("fluorobenzene", 2, 37, True), ("ala", 1, 64, False)])
def test_twisted_kernel_matches_oracle(name, n_steps, P, ref_norm):
    """cdw_langevin_params.twist (troubleshooting/csmc.py:460-623) against oracle/twisted.py, which is pinned to the reference's own outputs."""
    from coddiwomple_b200.engine import DeviceSystem
    from oracle import twisted
    from tests import gpu_util
    xml = ts.alanine_dipeptide_shaped_xml()[0] if name == "ala" else system_xml(name)
    o = XmlSystemOracle(xml)
    ds = DeviceSystem(xml)
    X = (cache_frames()[np.random.RandomState(0).randint(0, 151, P)] if name == "fluorobenzene" else ts.alanine_dipeptide_shaped_ensemble(P)).astype(np.float64)
    A = X.shape[1]
    r = np.random.RandomState(5)
    tw = twisted.pack_twist(40.0 * r.rand(A, 3), 3.0 * r.randn(A, 3), c=0.37, d=-0.11)
    lam = relative_lambdas(0.3)
    dt = 0.0005
    ref = twisted.twisted_ula(lambda xx: o.energy_forces(xx, lam), X, o.masses, o.constraints, dt, kT, n_steps, seed=7, gids=np.arange(5, 5 + P), step0=11,
                              twist=tw, reference_normalizer=ref_norm)
    flags = _lib.CDW_FLAG_ULA | (_lib.CDW_FLAG_TWIST_REFERENCE_NORMALIZER if ref_norm else 0)
    out = gpu_util.smc_propagate(ds, X, None, lam, dt, 0.0, kT, n_steps, seed=7, step0=11, gid0=5, flags=flags, aux=np.full(P, 0.25), cum=np.full(P, 1.5), twist=tw)
    assert np.abs(out["x"] - ref["x"]).max() < 2e-7
    np.testing.assert_allclose(out["u_last"], ref["u_last"] / kT, atol=2e-4)
    np.testing.assert_allclose(out["proposal"], ref["proposal_work"], rtol=1e-9, atol=3e-4)
    np.testing.assert_allclose(out["inc"], out["u_last"] + 0.25 + out["proposal"], rtol=1e-12)
    np.testing.assert_allclose(out["w"], 1.5 + out["inc"], rtol=1e-12)
    plain = gpu_util.smc_propagate(ds, X, None, lam, dt, 0.0, kT, n_steps, seed=7, step0=11, gid0=5, flags=_lib.CDW_FLAG_ULA, aux=np.full(P, 0.25), cum=np.full(P, 1.5))
    zero = gpu_util.smc_propagate(ds, X, None, lam, dt, 0.0, kT, n_steps, seed=7, step0=11, gid0=5, flags=flags, aux=np.full(P, 0.25), cum=np.full(P, 1.5),
                                  twist=np.zeros(6 * A + 2))
    assert np.abs(out["x"] - plain["x"]).max() > 1e-6
    assert np.abs(zero["x"] - plain["x"]).max() < 2e-7        # zero twist = the uncontrolled move (csmc.py:617-620)
    np.testing.assert_allclose(zero["inc"], plain["inc"], atol=2e-4)
    with pytest.raises(_lib.CdwError):                         # a twist without the Euler-Maruyama proposal is refused
        gpu_util.smc_propagate(ds, X, np.zeros_like(X), lam, dt, 1.0, kT, 1, seed=7, step0=11, twist=tw)


def test_twisted_engine_is_unbiased_and_a_good_twist_raises_the_effective_sample_size():
    """SMCEngine(proposal="ula", twist=TwistSequence(...)): the harmonic-oscillator protocol of the reference (Delta F = 1 kT, tests/utils.py:64-65)
    as plain SIS.  Any admissible twist leaves the estimate unbiased; a linear twist that pulls the proposal along the moving well raises the
    effective sample size of the final weights; the literal csmc.py:556 normaliser with A != 0 is visibly biased."""
    from coddiwomple_b200.engine import LangevinParameters, SMCEngine, TwistSequence
    prm = ts.harmonic_oscillator_parameters()
    P, T = 20000, 41
    x0 = np.random.RandomState(0).normal(0, prm["sigma"], (P, 1, 3)).astype(np.float32)
    seq = ts.harmonic_parameter_sequence(T)
    ramp = np.arange(1, T) / (T - 1)

    def run(twist):
        eng = SMCEngine(ts.harmonic_oscillator_xml(), P, seq, langevin=LangevinParameters(timestep=0.2), seed=1, resample_seed=2, proposal="ula",
                        floor_threshold=0.0, twist=twist)
        eng.anneal(x0)
        w = -eng.cumulative_works()
        return eng.free_energy(), float(np.exp(2 * np.logaddexp.reduce(w) - np.logaddexp.reduce(2 * w)) / P)

    def seq_twist(a, b, **kw):
        bb = np.zeros((T - 1, 1, 3)); bb[:, 0, 0] = b * 0.2 * ramp
        return TwistSequence(np.full((T - 1, 1, 3), a), bb, c=0.05 * np.arange(1, T), **kw)

    f0, n0 = run(None)
    f1, n1 = run(seq_twist(0.0, -12.0))
    f2, n2 = run(seq_twist(2.0, -10.0))
    f3, _ = run(seq_twist(2.0, -10.0, reference_normalizer=True))
    assert abs(f0 - 1.0) < 0.04 and abs(f1 - 1.0) < 0.04 and abs(f2 - 1.0) < 0.04, (f0, f1, f2)
    assert n1 > 1.25 * n0, (n0, n1)
    assert abs(f3 - 1.0) > 1.0, f3
    with pytest.raises(ValueError):
        SMCEngine(ts.harmonic_oscillator_xml(), P, seq, proposal="baoab", twist=seq_twist(0.0, -12.0))
    with pytest.raises(ValueError):   # 1 + 2 s a <= 0: not a Gaussian any more
        SMCEngine(ts.harmonic_oscillator_xml(), 16, seq, langevin=LangevinParameters(timestep=0.2), proposal="ula", twist=seq_twist(-1e4, 0.0))


@pytest.mark.parametrize("name,n_steps,P", [("fluorobenzene", 1, 100), ("fluorobenzene", 2, 37), ("ala", 1, 64)])
def test_dense_twisted_kernel_matches_oracle(name, n_steps, P):
    """cdw_langevin_params.dense_twist (a full matrix A_t, troubleshooting/csmc.py:460-623) against oracle.twisted.twisted_ula_dense."""
    from coddiwomple_b200.engine import DeviceSystem
    from oracle import twisted
    from tests import gpu_util
    xml = ts.alanine_dipeptide_shaped_xml()[0] if name == "ala" else system_xml(name)
    o = XmlSystemOracle(xml)
    ds = DeviceSystem(xml)
    X = (cache_frames()[np.random.RandomState(0).randint(0, 151, P)] if name == "fluorobenzene" else ts.alanine_dipeptide_shaped_ensemble(P)).astype(np.float64)
    A = X.shape[1]
    lam = relative_lambdas(0.3)
    dt = 0.0005
    r = np.random.RandomState(5)
    G = r.randn(3 * A, 3 * A) / np.sqrt(3 * A)
    s = np.repeat(kT * dt * dt / np.asarray(o.masses), 3)
    tab = twisted.dense_tables(15.0 * G @ G.T, 2.0 * r.randn(3 * A), 0.37, -0.11, s)
    ref = twisted.twisted_ula_dense(lambda xx: o.energy_forces(xx, lam), X, o.masses, o.constraints, dt, kT, n_steps, seed=7, gids=np.arange(5, 5 + P),
                                    step0=11, tables=tab)
    out = gpu_util.smc_propagate(ds, X, None, lam, dt, 0.0, kT, n_steps, seed=7, step0=11, gid0=5, flags=_lib.CDW_FLAG_ULA, aux=np.full(P, 0.25),
                                 cum=np.full(P, 1.5), twist=tab)
    assert np.abs(out["x"] - ref["x"]).max() < 2e-7
    np.testing.assert_allclose(out["u_last"], ref["u_last"] / kT, atol=2e-4)
    np.testing.assert_allclose(out["proposal"], ref["proposal_work"], rtol=1e-9, atol=3e-4)
    np.testing.assert_allclose(out["inc"], out["u_last"] + 0.25 + out["proposal"], rtol=1e-12)
    assert not out["nan"].any()
    with pytest.raises(_lib.CdwError):   # the dense descriptor excludes the reference-normaliser switch
        gpu_util.smc_propagate(ds, X, None, lam, dt, 0.0, kT, 1, seed=7, step0=11, flags=_lib.CDW_FLAG_ULA | _lib.CDW_FLAG_TWIST_REFERENCE_NORMALIZER, twist=tab)


def test_dense_twisted_engine_is_unbiased():
    """SMCEngine(proposal="ula", twist=TwistSequence.dense(...)) with a twist that couples x, y, z: Delta F = 1 kT (tests/utils.py:64-65); a twist whose
    precision matrix is not positive definite is refused."""
    from coddiwomple_b200.engine import LangevinParameters, SMCEngine, TwistSequence
    prm = ts.harmonic_oscillator_parameters()
    P, T = 20000, 41
    x0 = np.random.RandomState(0).normal(0, prm["sigma"], (P, 1, 3)).astype(np.float32)
    seq = ts.harmonic_parameter_sequence(T)
    A3 = np.array([[2.0, 0.8, -0.5], [0.8, 1.5, 0.3], [-0.5, 0.3, 1.0]])
    b = np.zeros((T - 1, 3)); b[:, 0] = -10.0 * 0.2 * np.arange(1, T) / (T - 1); b[:, 1] = 0.4; b[:, 2] = -0.3
    tw = TwistSequence.dense(np.repeat(A3[None], T - 1, 0), b, c=0.05 * np.arange(1, T))
    eng = SMCEngine(ts.harmonic_oscillator_xml(), P, seq, langevin=LangevinParameters(timestep=0.2), seed=1, resample_seed=2, proposal="ula",
                    floor_threshold=0.0, twist=tw)
    eng.anneal(x0)
    assert abs(eng.free_energy() - 1.0) < 0.04, eng.free_energy()
    assert int(eng.ensemble.nan_flags.sum()) == 0
    with pytest.raises(ValueError):
        SMCEngine(ts.harmonic_oscillator_xml(), 16, seq, langevin=LangevinParameters(timestep=0.2), proposal="ula",
                  twist=TwistSequence.dense(np.repeat(-1e4 * np.eye(3)[None], T - 1, 0), b))


def test_whole_twisted_smc_with_a_twisted_initial_draw():
    """twisted_smc end to end (troubleshooting/csmc.py:625-747): the initial particles come from the TWISTED Gaussian initial proposal with their
    log w_0 (coddiwomple_b200.csmc, csmc.py:336-458, 564-591), every move is twisted on the device.  Harmonic oscillator, pi_0 = N(0, sigma^2 I):
    Delta F = 1 kT (tests/utils.py:64-65) whatever the twists."""
    from coddiwomple_b200 import csmc
    from coddiwomple_b200.engine import LangevinParameters, SMCEngine, TwistSequence
    prm = ts.harmonic_oscillator_parameters()
    P, T = 20000, 41
    seq = ts.harmonic_parameter_sequence(T)
    sig2 = prm["sigma"] ** 2
    A0 = np.array([[6.0, 1.0, 0.0], [1.0, 4.0, -0.5], [0.0, -0.5, 3.0]])          # 1/nm^2, against a prior precision of 100 1/nm^2
    x0, log_w0 = csmc.twisted_gmm_initial_sample(P, [1.0], np.zeros((1, 3)), sig2 * np.eye(3)[None], A0, np.array([-3.0, 0.5, 0.0]), 0.2,
                                                 rng=np.random.default_rng(5))
    ramp = np.arange(1, T) / (T - 1)
    b = np.zeros((T - 1, 1, 3)); b[:, 0, 0] = -12.0 * 0.2 * ramp
    eng = SMCEngine(ts.harmonic_oscillator_xml(), P, seq, langevin=LangevinParameters(timestep=0.2), seed=1, resample_seed=2, proposal="ula",
                    floor_threshold=0.0, twist=TwistSequence(np.zeros((T - 1, 1, 3)), b, c=0.05 * np.arange(1, T)))
    eng.anneal(x0.reshape(P, 1, 3).astype(np.float32), initial_works=-log_w0)
    assert abs(eng.free_energy() - 1.0) < 0.04, eng.free_energy()
    plain = SMCEngine(ts.harmonic_oscillator_xml(), P, seq, langevin=LangevinParameters(timestep=0.2), seed=1, resample_seed=2, proposal="ula", floor_threshold=0.0)
    plain.anneal(x0.reshape(P, 1, 3).astype(np.float32))                            # the same draws WITHOUT their initial weights: biased
    assert abs(plain.free_energy() - 1.0) > 0.1, plain.free_energy()
