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
