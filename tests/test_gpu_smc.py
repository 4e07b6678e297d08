"""The fused annealing loop and the reference-facing API on the GPU: bookkeeping exactness against the oracle,
free-energy anchors of the reference, and the behavioural checks of coddiwomple/tests/test_openmm.py re-expressed."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from coddiwomple_b200 import _lib, testsystems as ts  # noqa: E402
from oracle import smc as oracle_smc, weights  # noqa: E402
from oracle.potential import XmlSystemOracle, kT_kj_per_mol  # noqa: E402
from tests.helpers import GOLDEN, cache_frames, generic_layer, system_xml  # noqa: E402

kT = kT_kj_per_mol()


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("the -m gpu tests need a CUDA device: there is no CPU fallback to test")


def _engine(xml, P, seq, **kw):
    from coddiwomple_b200.engine import SMCEngine
    return SMCEngine(xml, P, seq, **kw)


def test_every_iteration_obeys_the_reference_bookkeeping():
    """Record W, uniforms, ancestors and cum at every iteration of a GPU anneal and replay each iteration through the
    oracle's restatement of Resampler.resample (resamplers.py:61-133): ancestors bit-exact, works to 1e-12."""
    from coddiwomple_b200.engine import LangevinParameters
    frames = cache_frames()
    x0 = frames[np.random.RandomState(0).randint(0, len(frames), 512)]
    seq = ts.relative_alchemical_sequence(12)
    for kind, thr, always in (("multinomial", 0.9999, False), ("systematic", 0.5, True), ("multinomial", 0.5, False)):
        eng = _engine(system_xml("fluorobenzene"), 512, seq, langevin=LangevinParameters(timestep=0.002), resampler=kind, seed=3, resample_seed=4,
                      record_history=True, floor_threshold=thr, always_resample=always)
        eng.anneal(x0)
        cum_prev = np.zeros(512)
        stats = eng.stats.cpu().numpy()
        n_resampled = 0
        for h in eng.ensemble.history:
            t = h["t"]
            inc, W, cum, anc, u = (h[k].cpu().numpy() for k in ("inc", "W", "cum", "anc", "uniforms"))
            np.testing.assert_array_equal(W, cum_prev + inc)
            ref = weights.resample_step(cum_prev, inc, np.zeros(512), np.arange(512), uniforms=u, kind=kind, floor_threshold=thr, always=always)
            assert stats[t, _lib.STAT_NESS] == pytest.approx(weights.nESS(cum_prev, inc), rel=1e-12)
            assert bool(stats[t, _lib.STAT_RESAMPLE]) == ref["resampled"]
            np.testing.assert_array_equal(anc, ref["ancestors"])
            np.testing.assert_allclose(cum, ref["cum"], rtol=1e-12, atol=1e-12)
            n_resampled += ref["resampled"]
            cum_prev = cum
        assert (n_resampled >= 1 or thr == 0.5 and not always) and eng.resampled_iterations() == [h["t"] for h in eng.ensemble.history if stats[h["t"], _lib.STAT_RESAMPLE]]


def test_first_iterations_track_the_oracle_anneal():
    """Same inputs, same Philox counters: the GPU anneal and the oracle anneal agree on the incremental works of the
    first iterations (later ones drift apart at the 1e-6 level of the fp32 coordinates, then through resampling)."""
    frames = cache_frames()
    x0 = frames[np.random.RandomState(1).randint(0, len(frames), 128)]
    seq = ts.relative_alchemical_sequence(6)
    eng = _engine(system_xml("fluorobenzene"), 128, seq, seed=5, resample_seed=6, record_history=True)
    eng.anneal(x0)
    trace = []
    out = oracle_smc.anneal(XmlSystemOracle(system_xml("fluorobenzene")), x0, seq, seed=5, resample_seed=6, trace=trace)
    h = eng.ensemble.history
    np.testing.assert_allclose(h[0]["inc"].cpu().numpy(), trace[0]["inc"], rtol=1e-10)
    np.testing.assert_array_equal(h[0]["anc"].cpu().numpy(), trace[0]["ancestors"])
    if not trace[0]["resampled"]:
        np.testing.assert_allclose(h[1]["inc"].cpu().numpy(), trace[1]["inc"], atol=1e-4)
    assert abs(eng.free_energy() - out["free_energy"]) < 0.05


def test_config2_system_tracks_the_oracle_anneal():
    """The 22-atom AlanineDipeptideVacuum-shaped system of BASELINE configs[1] (team of 4 lanes, static-term cache on):
    the GPU anneal and oracle.smc.anneal (order of openmm/coddiwomple.py:312-331) run on the same particles with the
    same Philox counters: first-iteration incremental works at 1e-10, first ancestors bit-exact, second-iteration works
    at the fp32-coordinate level, and free energies equal within statistical error at matched P."""
    xml, _ = ts.alanine_dipeptide_shaped_xml()
    P = 384
    x0 = ts.alanine_dipeptide_shaped_ensemble(P, seed=7)
    seq = ts.relative_alchemical_sequence(16)
    o = XmlSystemOracle(xml)
    for kind in ("multinomial", "systematic"):
        eng = _engine(xml, P, seq, seed=5, resample_seed=6, resampler=kind, record_history=True)
        eng.anneal(x0)
        trace = []
        out = oracle_smc.anneal(o, x0, seq, seed=5, resample_seed=6, resampler=kind, trace=trace)
        h = eng.ensemble.history
        np.testing.assert_allclose(h[0]["inc"].cpu().numpy(), trace[0]["inc"], rtol=1e-10, atol=1e-10)
        np.testing.assert_array_equal(h[0]["anc"].cpu().numpy(), trace[0]["ancestors"])
        if not trace[0]["resampled"]:
            np.testing.assert_allclose(h[1]["inc"].cpu().numpy(), trace[1]["inc"], atol=2e-4)
        # the statistical error of the estimate at this P is ~0.1 kT; both runs share particles, noise and uniforms
        assert abs(eng.free_energy() - out["free_energy"]) < 0.1, (eng.free_energy(), out["free_energy"])
        assert int(eng.ensemble.nan_flags.sum()) == 0


def test_harmonic_oscillator_free_energy_is_one_kT():
    """G6: x0 0 -> 2 sigma, U0 0 -> 1 kT gives Delta F = 1 kT exactly (coddiwomple/tests/utils.py:63-67); config 1."""
    from coddiwomple_b200.engine import LangevinParameters
    prm = ts.harmonic_oscillator_parameters()
    x0 = np.random.RandomState(0).normal(0, prm["sigma"], (20000, 1, 3)).astype(np.float32)
    for kind in ("multinomial", "systematic"):
        eng = _engine(ts.harmonic_oscillator_xml(), 20000, ts.harmonic_parameter_sequence(21),
                      langevin=LangevinParameters(timestep=prm["timestep"], collision_rate=prm["collision_rate"], n_steps=5), resampler=kind, seed=1, resample_seed=2)
        eng.anneal(x0)
        assert abs(eng.free_energy() - 1.0) < 0.03, eng.free_energy()
        assert int(eng.ensemble.nan_flags.sum()) == 0


def test_fluorobenzene_free_energy_band_of_the_reference():
    """G5: 1.33 +- 0.2 kT for benzene -> fluorobenzene, 10 lambdas, 1 step per lambda (legacy_tests.py:90-94)."""
    frames = cache_frames()
    x0 = frames[np.random.RandomState(3).randint(0, len(frames), 4000)]
    eng = _engine(system_xml("fluorobenzene"), 4000, ts.relative_alchemical_sequence(10), seed=11, resample_seed=12)
    eng.anneal(x0)
    assert abs(eng.free_energy() - 1.33) < 0.2, eng.free_energy()
    x = eng.ensemble.positions.double().cpu().numpy()
    o = XmlSystemOracle(system_xml("fluorobenzene"))
    for (i, j, d) in o.constraints:   # particles stay on the constraint manifold through propagation and resampling
        assert np.abs(np.linalg.norm(x[:, i] - x[:, j], axis=-1) / d - 1).max() < 5e-6
    # aux carried with the resampled state is -u_T(x_T) (openmm/coddiwomple.py:326-327)
    u = o.reduced_potential(x, ts.relative_alchemical_sequence(10)[-1])
    np.testing.assert_allclose(eng.ensemble.auxiliary_work.cpu().numpy(), -u, rtol=1e-9)


def test_mcmc_smc_resampler_signature_and_result():
    """coddiwomple/tests/legacy_tests.py:52-95 re-expressed: same call, same assertions."""
    from coddiwomple_b200.openmm.coddiwomple import mcmc_smc_resampler
    import os
    lambda_sequence = ts.relative_alchemical_sequence(10)
    np.random.seed(0)
    particles = mcmc_smc_resampler(system=system_xml("fluorobenzene"), endstate_cache_filename=os.path.join(GOLDEN, "fluorobenzene_cache.npy"),
                                   directory_name=None, trajectory_prefix=None, md_topology=None, number_of_particles=10,
                                   parameter_sequence=lambda_sequence)
    assert len(particles) == 10
    assert all(particle.iteration == 9 for particle in particles)
    assert all(np.isfinite(particle.cumulative_work) for particle in particles)
    np.random.seed(1)
    particles = mcmc_smc_resampler(system_xml("fluorobenzene"), os.path.join(GOLDEN, "fluorobenzene_cache.npy"), None, None, None, 3000, lambda_sequence)
    average_free_energy = np.mean([particle.cumulative_work for particle in particles[:50]])  # all equal after a final resample
    fe = particles.engine.free_energy()
    assert (1.33 - 0.2) / 1.33 < fe / 1.33 < (1.33 + 0.2) / 1.33
    assert np.isfinite(average_free_energy)
    p = particles[0]
    assert p.state.positions.shape == (13, 3) and p.ancestry[0] == 0 and len(p.proposal_works) == 10
    # per-iteration histories like the reference's Particle (particles.py:38-47; appended at resamplers.py:115,129,196-200):
    # one incremental work and one root label per iteration; the works add up to the cumulative work; labels are roots
    eng = particles.engine
    stats = eng.stats.cpu().numpy()
    resampled = [t for t in range(1, 10) if stats[t, _lib.STAT_RESAMPLE]]
    for i in (0, 17, 2999):
        p = particles[i]
        assert len(p.incremental_works) == 9 and len(p.ancestry) == 10 and p.ancestry[0] == i
        assert sum(p.incremental_works) == pytest.approx(p.cumulative_work, abs=1e-9)
        assert all(0 <= a < 3000 for a in p.ancestry)
        for t in range(1, 10):   # the root label only changes at iterations that resampled
            if t not in resampled:
                assert p.ancestry[t] == p.ancestry[t - 1]
    assert particles[5].ancestry[-1] == int(eng.ensemble.labels[5])
    if resampled:   # right after a resampling every particle carries the ensemble normaliser (resamplers.py:107,129)
        t = resampled[0]
        assert len({round(sum(particles[i].incremental_works[:t]), 12) for i in (0, 17, 2999)}) == 1
    np.random.seed(2)
    lean = mcmc_smc_resampler(system_xml("fluorobenzene"), os.path.join(GOLDEN, "fluorobenzene_cache.npy"), None, None, None, 64, lambda_sequence,
                              record_histories=False)
    assert np.isfinite(lean[0].cumulative_work)
    with pytest.raises(RuntimeError):
        len(lean[0].incremental_works)
    with pytest.raises(RuntimeError):
        lean[0].ancestry[-1]


def test_OpenMMPDFState_contract():
    """coddiwomple/tests/test_openmm.py:24-59."""
    from coddiwomple_b200.openmm.states import HarmonicAlchemicalState, OpenMMParticleState, OpenMMPDFState
    pdf_state = OpenMMPDFState(system=ts.harmonic_oscillator_xml(), alchemical_composability=HarmonicAlchemicalState, temperature=300.0, pressure=None)
    assert set(pdf_state.get_parameters()) == {f"{ts.HO_PREFIX}_x0", f"{ts.HO_PREFIX}_U0"}
    pdf_state.set_parameters({f"{ts.HO_PREFIX}_x0": 1.0, f"{ts.HO_PREFIX}_U0": 1.0})
    assert pdf_state.get_parameters() == {f"{ts.HO_PREFIX}_x0": 1.0, f"{ts.HO_PREFIX}_U0": 1.0}
    with pytest.raises(AssertionError):
        pdf_state.set_parameters({f"{ts.HO_PREFIX}_x0": 1.0})
    state = OpenMMParticleState(positions=np.zeros((1, 3)))
    K = ts.harmonic_oscillator_parameters()["K"]
    assert pdf_state.reduced_potential(state) == pytest.approx((0.5 * K * 1.0 + 1.0) / kT, rel=1e-12)
    with pytest.raises(NotImplementedError):
        OpenMMPDFState(system=ts.harmonic_oscillator_xml(), pressure=1.0)


def test_OMMLI_and_OMMBIP_contract():
    """coddiwomple/tests/test_openmm.py:157-288 re-expressed on the harmonic test system."""
    from coddiwomple_b200.openmm.integrators import OMMLI
    from coddiwomple_b200.openmm.propagators import OMMBIP
    from coddiwomple_b200.openmm.states import HarmonicAlchemicalState, OpenMMParticleState, OpenMMPDFState
    prm = ts.harmonic_oscillator_parameters()
    pdf_state = OpenMMPDFState(system=ts.harmonic_oscillator_xml(), alchemical_composability=HarmonicAlchemicalState, temperature=300.0, pressure=None)
    integrator = OMMLI(temperature=300.0, collision_rate=prm["collision_rate"], timestep=prm["timestep"])
    propagator = OMMBIP(openmm_pdf_state=pdf_state, integrator=integrator, seed=4)
    assert propagator.pdf_state is pdf_state                       # :260 "VERY important for SMC"
    particle_state = OpenMMParticleState(positions=np.zeros((1, 3)))
    # null application (:264-268)
    prior = pdf_state.reduced_potential(particle_state)
    _, proposal_work = propagator.apply(particle_state, n_steps=0)
    assert proposal_work == 0. and pdf_state.reduced_potential(particle_state) == prior
    assert propagator.context_reduced_potential(particle_state) == pytest.approx(prior)
    # context update internals (:271-281)
    parameters = pdf_state.get_parameters()
    parameters[f"{ts.HO_PREFIX}_U0"] = 1.
    pdf_state.set_parameters(parameters)
    propagator.apply(particle_state, n_steps=0, apply_pdf_to_context=False)
    assert propagator._get_context_parameters()[f"{ts.HO_PREFIX}_U0"] == 0.
    assert propagator.context_reduced_potential(particle_state) == pytest.approx(prior)
    propagator.apply(particle_state, n_steps=0, apply_pdf_to_context=True)
    assert propagator._get_context_parameters()[f"{ts.HO_PREFIX}_U0"] == 1.
    assert propagator.context_reduced_potential(particle_state) == pytest.approx(prior + 1.0 / kT, rel=1e-12)
    parameters[f"{ts.HO_PREFIX}_U0"] = 0.
    pdf_state.set_parameters(parameters)
    propagator.apply(particle_state, n_steps=0, apply_pdf_to_context=True)
    # equilibrium sampling (:195-221): 100 x 20 steps with velocity re-randomisation
    reduced_pe = []
    for _ in range(100):
        particle_state, proposal_work = propagator.apply(particle_state, n_steps=20, reset_integrator=False, apply_pdf_to_context=False, randomize_velocities=True)
        assert proposal_work == 0.
        assert integrator._get_energy_with_units('proposal_work', dimensionless=True) != 0.
        assert integrator._get_energy_with_units('shadow_work', dimensionless=True) != 0.
        reduced_pe.append(pdf_state.reduced_potential(particle_state))
    tol = 6 * np.sqrt(1.5)
    assert 1.5 - tol < np.mean(reduced_pe) < 1.5 + tol
    assert propagator.apply(particle_state, n_steps=1, returnable_key="proposal")[1] == integrator.get_proposal_work()
    with pytest.raises(Exception):
        propagator.apply(particle_state, n_steps=1, returnable_key="nonexistent")
    integrator.reset()
    assert integrator._get_energy_with_units('proposal_work', dimensionless=True) == 0.
    assert integrator._get_energy_with_units('shadow_work', dimensionless=True) == 0.
    # stability (:288)
    propagator.apply(particle_state, n_steps=1000, reset_integrator=True, apply_pdf_to_context=True, randomize_velocities=True)
    assert np.isfinite(particle_state.positions).all()
    with pytest.raises(NotImplementedError):
        OMMLI(splitting="V0 R V1")      # force-group (multiple-time-step) splittings: refused, as in the reference (integrators.py:412)
    with pytest.raises(ValueError):
        OMMLI(splitting="V { R } V")    # shadow-work generating steps outside the Metropolised block (integrators.py:248-276)
    with pytest.raises(ValueError):
        OMMLI(splitting="{ V R { O } R V }")


def test_metropolised_integrator_samples_the_right_distribution():
    """OMMLI with a Metropolised block (openmm/integrators.py:427-444): GHMC "O { V R V } O" at a time step where plain Langevin
    dynamics is visibly biased still samples <u> = 3/2 on the harmonic oscillator; trials = accepted + rejected."""
    from coddiwomple_b200.openmm.integrators import OMMLI
    from coddiwomple_b200.openmm.propagators import OMMBIP
    from coddiwomple_b200.openmm.states import HarmonicAlchemicalState, OpenMMPDFState
    prm = ts.harmonic_oscillator_parameters()
    pdf_state = OpenMMPDFState(system=ts.harmonic_oscillator_xml(), alchemical_composability=HarmonicAlchemicalState, pressure=None)
    dt = 20.0 * prm["timestep"]     # omega dt = 1 ("period" of tests/utils.py:55-60 is 1 / omega): velocity Verlet samples x with variance sigma^2 / (1 - (omega dt)^2 / 4)
    means = {}
    for splitting in ("O { V R V } O", "O V R V O"):
        integ = OMMLI(collision_rate=prm["collision_rate"], timestep=dt, splitting=splitting)
        prop = OMMBIP(pdf_state, integ, seed=7)
        pos = torch.zeros(20000, 1, 4, device="cuda")
        vel = torch.zeros_like(pos)
        for _ in range(4):
            prop.apply_batch(pos, vel, n_steps=50, randomize_velocities=True)
        means[splitting] = float(pdf_state.reduced_potential_batch(pos).mean())
        if "{" in splitting:
            assert integ.ntrials == 4 * 50 * 20000 and integ.naccept + integ.nreject == integ.ntrials
            assert 0.3 < integ.acceptance_rate < 0.999
    assert abs(means["O { V R V } O"] - 1.5) < 0.04, means
    assert abs(means["O V R V O"] - 2.0) < 0.06, means   # <u> = 1.5 / (1 - 1/4): the bias the Metropolis step removes


def test_OMMBIP_restart_attempts():
    """openmm/propagators.py:134-190: a move that ends in a non-finite energy is retried n_restart_attempts times (fresh
    noise); only the failed replicas are integrated again; what still fails stays at its pre-move state."""
    from coddiwomple_b200.engine import positions_to_rows
    from coddiwomple_b200.openmm.integrators import OMMLI
    from coddiwomple_b200.openmm.propagators import OMMBIP
    from coddiwomple_b200.openmm.states import OpenMMPDFState, RelativeAlchemicalState
    X = cache_frames()[:48].astype(np.float64)
    X[7, 12] = X[7, 5]            # two atoms on top of each other: the energy is not finite whatever the noise
    runs = {}
    for attempts in (0, 2):
        pdf_state = OpenMMPDFState(system_xml("fluorobenzene"), RelativeAlchemicalState, temperature=300.0, pressure=None)
        prop = OMMBIP(pdf_state, OMMLI(), n_restart_attempts=attempts, seed=3)
        pos, vel = positions_to_rows(X), torch.zeros(48, 13, 4, device="cuda")
        out = prop.apply_batch(pos, vel, n_steps=2, randomize_velocities=True, apply_pdf_to_context=True)
        assert out["attempts"] == attempts and int(out["nan"].sum()) == 1 and int(out["nan"][7]) == 1
        np.testing.assert_array_equal(pos[7, :, :3].cpu().numpy(), X[7].astype(np.float32))   # reverted
        runs[attempts] = pos.cpu().numpy()
    np.testing.assert_array_equal(runs[0], runs[2])   # the 47 good replicas were integrated exactly once


def test_batched_equilibrium_statistics_many_particles():
    """The same G6 check on 20k replicas at once: <u> = 3/2 and std = sqrt(3/2) to ~1 %."""
    from coddiwomple_b200.openmm.integrators import OMMLI
    from coddiwomple_b200.openmm.propagators import OMMBIP
    from coddiwomple_b200.openmm.states import HarmonicAlchemicalState, OpenMMPDFState
    prm = ts.harmonic_oscillator_parameters()
    pdf_state = OpenMMPDFState(system=ts.harmonic_oscillator_xml(), alchemical_composability=HarmonicAlchemicalState, pressure=None)
    prop = OMMBIP(pdf_state, OMMLI(collision_rate=prm["collision_rate"], timestep=prm["timestep"]), seed=7)
    pos = torch.zeros(20000, 1, 4, device="cuda")
    vel = torch.zeros_like(pos)
    prop.apply_batch(pos, vel, n_steps=400, randomize_velocities=True)
    u = pdf_state.reduced_potential_batch(pos).cpu().numpy()
    assert abs(u.mean() - 1.5) < 0.03 and abs(u.std() - np.sqrt(1.5)) < 0.04


@pytest.mark.parametrize("case", [1, 2, 4])
def test_Resampler_on_python_particles_matches_the_reference(case):
    """`MultinomialResampler.resample` on a plain list of `Particle` objects (the reference's own calling convention),
    under the same global numpy seed the reference consumed: ancestors, labels and aux identical, works to 1e-12."""
    from coddiwomple_b200.particles import Particle
    from coddiwomple_b200.resamplers import MultinomialResampler
    from coddiwomple_b200.utils import nESS, vanilla_floor_threshold
    c = generic_layer()["cases"][case]
    P = c["P"]
    particles = [Particle(index=i) for i in range(P)]
    for p, cw, a in zip(particles, c["cumulative_works_in"], c["auxiliary_works_in"]):
        p.update_work(cw)
        p.update_auxiliary_work(a)
        p.update_state(("state", p.ancestry[0]))
    assert nESS(particles, np.array(c["incremental_works"])) == pytest.approx(c["nESS"], rel=1e-12)
    np.random.seed(1000 + c["seed"])
    res = MultinomialResampler()
    res.resample(particles, np.array(c["incremental_works"]), observable=nESS, threshold=vanilla_floor_threshold)
    assert res.resampling_logger == c["resampling_logger"]
    np.testing.assert_array_equal([p.state[1] for p in particles], c["after_thresholded"]["state_label"])
    np.testing.assert_array_equal([p.ancestry[-1] for p in particles], c["after_thresholded"]["ancestry_last"])
    np.testing.assert_array_equal([p.auxiliary_work for p in particles], c["after_thresholded"]["auxiliary_work"])
    np.testing.assert_allclose([p.cumulative_work for p in particles], c["after_thresholded"]["cumulative_work"], rtol=1e-12, atol=1e-12)


def test_annealed_importance_sampling_driver_and_propagator():
    """AIS mode (coddiwomple/tests/legacy_tests.py:24-50 re-expressed): the state-work record has the
    reference's shape, the works equal the oracle's resampling-free anneal at the first steps, and the Jarzynski estimate of
    the harmonic test system is 1 kT."""
    import os
    from coddiwomple_b200.openmm.coddiwomple import annealed_importance_sampling
    from coddiwomple_b200.openmm.integrators import OMMLIAIS
    from coddiwomple_b200.openmm.propagators import OMMAISP
    from coddiwomple_b200.openmm.states import HarmonicAlchemicalState, OpenMMParticleState, OpenMMPDFState, RelativeAlchemicalState
    x = 'fractional_iteration'
    alchemical_functions = {'lambda_sterics_core': x, 'lambda_electrostatics_core': x,
                            'lambda_sterics_insert': f"select(step({x} - 0.5), 1.0, 2.0 * {x})",
                            'lambda_sterics_delete': f"select(step({x} - 0.5), 2.0 * ({x} - 0.5), 0.0)",
                            'lambda_electrostatics_insert': f"select(step({x} - 0.5), 2.0 * ({x} - 0.5), 0.0)",
                            'lambda_electrostatics_delete': f"select(step({x} - 0.5), 1.0, 2.0 * {x})",
                            'lambda_bonds': x, 'lambda_angles': x, 'lambda_torsions': x}
    np.random.seed(3)
    works = annealed_importance_sampling(system_xml("fluorobenzene"), os.path.join(GOLDEN, "fluorobenzene_cache.npy"), None, None, None, alchemical_functions,
                                         number_of_applications=512, steps_per_application=10, endstate_parameters=0.0, record_state_work_interval=1, seed=5)
    assert sorted(works) == list(range(512)) and all(len(w) == 11 and w[0] == 0.0 for w in works.values())
    final = np.array([w[-1] for w in works.values()])
    fe = -(np.log(np.mean(np.exp(-(final - final.min())))) - final.min())
    assert abs(fe - 1.33) < 0.25, fe                                # same free energy as the SMC route (legacy_tests.py:90-94)
    # first increment is deterministic: u_1(x_0) - u_0(x_0) on the chosen frames
    np.random.seed(3)
    frames = cache_frames()
    chosen = np.random.choice(range(len(frames)), 512)
    o = XmlSystemOracle(system_xml("fluorobenzene"))
    seq = ts.relative_alchemical_sequence(11)
    dw = o.reduced_potential(frames[chosen].astype(np.float64), seq[1]) - o.reduced_potential(frames[chosen].astype(np.float64), seq[0])
    np.testing.assert_allclose([w[1] for w in works.values()], dw, rtol=1e-8, atol=1e-9)
    # single-state propagator API on the harmonic oscillator + Jarzynski
    prm = ts.harmonic_oscillator_parameters()
    harmonic = {f"{ts.HO_PREFIX}_x0": f"(1-{x})*0.0 + {x}*{2 * prm['sigma']!r}", f"{ts.HO_PREFIX}_U0": f"(1-{x})*0.0 + {x}*{prm['kT']!r}"}
    pdf_state = OpenMMPDFState(ts.harmonic_oscillator_xml(), alchemical_composability=HarmonicAlchemicalState, pressure=None)
    integrator = OMMLIAIS(harmonic, 200, temperature=300.0, collision_rate=prm["collision_rate"], timestep=prm["timestep"])
    prop = OMMAISP(pdf_state, integrator, record_state_work_interval=50, reassign_velocities=True, seed=2)
    state = OpenMMParticleState(positions=np.zeros((1, 3)))
    _, pw = prop.apply(state, n_steps=200, reset_integrator=True, apply_pdf_to_context=True)
    assert pw == 0. and len(prop.state_works[0]) == 5 and prop.state_works[0][0] == 0.0
    assert pdf_state.get_parameters()[f"{ts.HO_PREFIX}_U0"] == pytest.approx(prm["kT"])
    pdf_state.set_parameters({k: 0.0 for k in pdf_state.get_parameters()})
    x0 = np.random.RandomState(0).normal(0, prm["sigma"], (20000, 1, 3)).astype(np.float32)
    w, _ = prop.apply_batch(x0, n_steps=200)
    assert w.shape == (20000, 5)
    fe = -np.log(np.mean(np.exp(-w[:, -1])))   # work std ~0.9 kT at 200 steps: the Jarzynski estimate is good to ~0.01 kT
    assert abs(fe - 1.0) < 0.04, fe
    with pytest.raises(AssertionError):
        OMMAISP(OpenMMPDFState(system_xml("fluorobenzene"), alchemical_composability=RelativeAlchemicalState, pressure=None), integrator)


def test_prebuilt_protocol_tables_are_bit_identical():
    """cdw_system_register_protocol only changes WHERE the effective term tables come from (one bulk copy of a prebuilt
    image instead of the per-CTA prologue): every output of the anneal is bitwise the same."""
    frames = cache_frames()
    x0 = frames[np.random.RandomState(2).randint(0, len(frames), 700)]
    seq = ts.relative_alchemical_sequence(8)
    runs = []
    for clear in (False, True):
        eng = _engine(system_xml("fluorobenzene"), 700, seq, seed=5, resample_seed=6)
        if clear:
            _lib.check(eng.system.lib.cdw_system_clear_protocol(eng.system.handle), "cdw_system_clear_protocol")
        eng.anneal(x0)
        runs.append((eng.ensemble.cum.cpu().numpy().copy(), eng.ensemble.positions.cpu().numpy().copy(), eng.ensemble.auxiliary_work.cpu().numpy().copy()))
    for a, b in zip(*runs):
        np.testing.assert_array_equal(a, b)


def test_config3_full_size_perses_vacuum_1e5_particles():
    """BASELINE configs[2] at full size (10^5 particles, fluorobenzene hybrid, the reference's 10-lambda protocol), checked
    through size-independent properties: the G5 free-energy band, the constraint manifold, aux = -u_T(x_T) against the
    oracle on a sample, root labels, and the weight bookkeeping of the last iteration."""
    frames = cache_frames()
    P = 100_000
    x0 = frames[np.random.RandomState(3).randint(0, len(frames), P)]
    seq = ts.relative_alchemical_sequence(10)
    eng = _engine(system_xml("fluorobenzene"), P, seq, seed=21, resample_seed=22)
    eng.anneal(x0)
    assert abs(eng.free_energy() - 1.33) < 0.2, eng.free_energy()
    assert int(eng.ensemble.nan_flags.sum()) == 0
    x = eng.ensemble.positions.double().cpu().numpy()
    o = XmlSystemOracle(system_xml("fluorobenzene"))
    for (i, j, d) in o.constraints:
        assert np.abs(np.linalg.norm(x[:, i] - x[:, j], axis=-1) / d - 1).max() < 5e-6
    sample = np.random.RandomState(0).randint(0, P, 256)
    np.testing.assert_allclose(eng.ensemble.auxiliary_work.cpu().numpy()[sample], -o.reduced_potential(x[sample], seq[-1]), rtol=1e-9)
    labels = eng.ensemble.labels.cpu().numpy()
    assert labels.min() >= 0 and labels.max() < P
    stats = eng.stats.cpu().numpy()
    ness = stats[1:, _lib.STAT_NESS]
    assert np.all((ness > 0) & (ness <= 1 + 1e-12))
    cum = eng.ensemble.cum.cpu().numpy()
    assert np.isfinite(cum).all()
    if stats[len(seq) - 1, _lib.STAT_RESAMPLE]:      # after a resampling every particle carries the same cumulative work
        assert np.ptp(cum) == 0.0


def test_config2_full_size_graph_replay_equals_eager_loop():
    """BASELINE configs[1] (the bench workload: 10^4 alanine-dipeptide-shaped replicas, 100 lambdas): the CUDA-graph
    replay and the eager loop give bitwise the same anneal; no NaN; constraints hold; nESS and labels are sane."""
    from coddiwomple_b200.engine import LangevinParameters
    xml, _ = ts.alanine_dipeptide_shaped_xml()
    P = 10_000
    x0 = ts.alanine_dipeptide_shaped_ensemble(P, seed=2)
    seq = ts.relative_alchemical_sequence(100)
    lang = LangevinParameters(temperature=300.0, collision_rate=1.0, timestep=0.002, n_steps=1)
    eager = _engine(xml, P, seq, langevin=lang, seed=0, resample_seed=1)
    eager.anneal(x0)
    graph = _engine(xml, P, seq, langevin=lang, seed=0, resample_seed=1)
    from coddiwomple_b200.engine import positions_to_rows
    rows = positions_to_rows(x0)
    graph.load(rows)
    graph.capture()
    graph.anneal(rows)
    np.testing.assert_array_equal(eager.ensemble.cum.cpu().numpy(), graph.ensemble.cum.cpu().numpy())
    np.testing.assert_array_equal(eager.ensemble.positions.cpu().numpy(), graph.ensemble.positions.cpu().numpy())
    assert eager.free_energy() == graph.free_energy() and np.isfinite(eager.free_energy())
    assert int(eager.ensemble.nan_flags.sum()) == 0
    x = eager.ensemble.positions.double().cpu().numpy()
    o = XmlSystemOracle(xml)
    for (i, j, d) in o.constraints:
        assert np.abs(np.linalg.norm(x[:, i] - x[:, j], axis=-1) / d - 1).max() < 5e-6
    assert len(eager.resampled_iterations()) >= 1
    labels = eager.ensemble.labels.cpu().numpy()
    assert labels.min() >= 0 and labels.max() < P


@pytest.mark.parametrize("kind,P,n_steps", [("multinomial", 3000, 1), ("systematic", 10000, 1), ("multinomial", 12001, 2), ("multinomial", 40000, 1)])
def test_lookahead_loop_equals_the_sequential_loop(kind, P, n_steps):
    """The look-ahead loop (cdw_smc_lookahead: the evaluation that opens iteration t + 1 closes launch t, the resampling runs on a
    side stream beside the next launch) and the sequential loop (propagate, then resample, openmm/coddiwomple.py:314-328 in
    order) give bitwise the same anneal: weights, statistics, ancestors, states."""
    from coddiwomple_b200.engine import LangevinParameters
    xml, _ = ts.alanine_dipeptide_shaped_xml()
    x0 = ts.alanine_dipeptide_shaped_ensemble(P, seed=4)
    seq = ts.relative_alchemical_sequence(9)
    runs = []
    for lookahead in (True, False):
        eng = _engine(xml, P, seq, langevin=LangevinParameters(n_steps=n_steps), resampler=kind, seed=8, resample_seed=9, floor_threshold=0.9999,
                      lookahead=lookahead, record_lineage=True)
        eng.initialize(x0)
        for t in range(1, eng.T):
            eng.step(t)
        eng.finish()
        e = eng.ensemble
        lin = eng.lineage()
        runs.append([v.cpu().numpy().copy() for v in (e.cum, eng.stats, e.anc, e.pos[e.cur], e.vel[e.cur], e.aux[e.cur], e.label[e.cur], e.inc, e.W)]
                    + [lin["cum"], lin["labels"]])
        assert len(eng.resampled_iterations()) >= 2
        assert int(e.nan_flags.sum()) == 0
    for k, (a, b) in enumerate(zip(*runs)):
        np.testing.assert_array_equal(a, b, err_msg=str(k))


@pytest.mark.parametrize("team", [1, 2, 8])
def test_other_team_sizes_run_both_loops(team, tmp_path):
    """CDW_TEAM = 1, 2 (32- / 64-thread CTAs) and 8 (256-thread CTAs): for every launch shape the look-ahead loop and the
    sequential loop agree bit for bit (the team size only changes the summation order inside a replica, for both alike)."""
    import os
    import subprocess
    import sys
    probe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "team_probe.py")
    outs = []
    for mode in ("lookahead", "sequential"):
        env = dict(os.environ, CDW_TEAM=str(team))
        out = str(tmp_path / f"team{team}_{mode}.npz")
        subprocess.run([sys.executable, probe, out, mode], env=env, check=True, timeout=600)
        outs.append(np.load(out))
    for k in outs[0].files:
        np.testing.assert_array_equal(outs[0][k], outs[1][k], err_msg=k)
    stats = outs[0]["stats"]
    assert stats[1:, _lib.STAT_RESAMPLE].sum() >= 2 and np.all(stats[1:, _lib.STAT_NESS] > 0)
    assert outs[0]["anc"].min() >= 0 and outs[0]["anc"].max() < 3000


@pytest.mark.parametrize("P", [1, 2, 5, 33])
def test_tiny_ensembles(P):
    """Edge sizes: fewer particles than one warp team / one CTA; ancestors stay in range, a single particle keeps weight 1."""
    frames = cache_frames()
    x0 = frames[np.random.RandomState(9).randint(0, len(frames), P)]
    seq = ts.relative_alchemical_sequence(5)
    eng = _engine(system_xml("fluorobenzene"), P, seq, seed=3, resample_seed=4, always_resample=True, record_history=True)
    eng.anneal(x0)
    for h in eng.ensemble.history:
        anc = h["anc"].cpu().numpy()
        assert anc.min() >= 0 and anc.max() < P
        ref = weights.resample_step(h["W"].cpu().numpy() - h["inc"].cpu().numpy(), h["inc"].cpu().numpy(), np.zeros(P), np.arange(P),
                                    uniforms=h["uniforms"].cpu().numpy(), kind="multinomial", always=True)
        np.testing.assert_array_equal(anc, ref["ancestors"])
    assert np.isfinite(eng.free_energy())
    stats = eng.stats.cpu().numpy()
    assert np.all(stats[1:, _lib.STAT_NESS] > 0) and np.all(stats[1:, _lib.STAT_NESS] <= 1 + 1e-12)
    if P == 1:
        np.testing.assert_allclose(stats[1:, _lib.STAT_NESS], 1.0, rtol=1e-12)


def test_SMC_with_the_callers_own_observable_threshold_resampler_and_termination():
    """`SMC()` with pluggable pieces, as the reference's interfaces allow (resamplers.py:61-133, 202-216; utils.py:41-72;
    distribution_factories.py:134-148): a custom observable / threshold pair, a custom `_resample`, and `termination_parameters` that end
    the loop early.  The particles stay on the device; the custom rule that restates the stock one reproduces the device-decided run bit for bit."""
    from coddiwomple_b200.distribution_factories import ProposalFactory, TargetFactory
    from coddiwomple_b200.openmm.integrators import OMMLI
    from coddiwomple_b200.openmm.propagators import OMMBIP
    from coddiwomple_b200.openmm.states import OpenMMParticleState, OpenMMPDFState, RelativeAlchemicalState
    from coddiwomple_b200.resamplers import MultinomialResampler, Resampler
    from coddiwomple_b200.smc import SMC
    from coddiwomple_b200.utils import nESS, vanilla_floor_threshold
    seq = ts.relative_alchemical_sequence(8)
    P = 512
    x0 = cache_frames()[np.random.RandomState(4).randint(0, 151, P)]
    pdf = OpenMMPDFState(system=system_xml("fluorobenzene"), alchemical_composability=RelativeAlchemicalState, pressure=None)

    def run(observable, threshold, resampler, term=None, **kw):
        prop = OMMBIP(pdf, OMMLI(temperature=300.0, collision_rate=1.0, timestep=0.002))
        return SMC(TargetFactory(pdf, list(seq), termination_parameters=term), ProposalFactory(list(seq), prop), P, None, resampler,
                   initial_positions=x0, observable=observable, threshold=threshold, seed=11,
                   state_factory=lambda x, v: OpenMMParticleState(positions=x, velocities=v), **kw)

    stock = run(nESS, vanilla_floor_threshold, MultinomialResampler(), floor_threshold=0.9999)
    calls = []

    def my_observable(particles, incremental_works, **kwargs):      # the reference's nESS (utils.py:41-64), restated in numpy
        w = -(np.array([p.cumulative_work for p in particles]) + np.asarray(incremental_works))
        w = np.exp(w - np.logaddexp.reduce(w))
        calls.append(len(particles))
        return 1.0 / (len(w) * np.sum(w ** 2))

    custom = run(my_observable, lambda value, **kw: not (value > 0.9999), MultinomialResampler())
    assert len(calls) == 7 and set(calls) == {P}
    assert custom.engine.resampled_iterations() == stock.engine.resampled_iterations() and len(stock.engine.resampled_iterations()) > 0
    np.testing.assert_array_equal(custom.cumulative_works, stock.cumulative_works)                 # same decisions, same kernels: same bits
    np.testing.assert_array_equal(custom.engine.ensemble.positions.cpu().numpy(), stock.engine.ensemble.positions.cpu().numpy())

    class EveryOther(Resampler):                                     # a resampler of the caller's own: particle i takes over particle i - (i % 2)
        def _resample(self, cumulative_works, num_resamples, **kwargs):
            self.seen = np.array(cumulative_works, dtype=np.float64)
            return [i - (i % 2) for i in range(num_resamples)]

    mine = EveryOther()
    out = run(None, None, mine)                                      # observable None: resample at every iteration (resamplers.py:90-92)
    labels = out.engine.ensemble.labels.cpu().numpy()
    np.testing.assert_array_equal(labels, np.arange(P) - (np.arange(P) % 2))
    mean = -(np.logaddexp.reduce(-mine.seen) - np.log(P))            # every particle carries the normaliser of the last resampling
    np.testing.assert_allclose(out.cumulative_works, mean, rtol=1e-12)
    assert out.engine.resampled_iterations() == list(range(1, 8)) and mine.resampling_logger == list(range(1, 8))
    np.testing.assert_array_equal(out[3].state.positions, out[2].state.positions)
    assert out[3].ancestry[-1] == 2 and len(out[3].incremental_works) == 7

    early = run(nESS, vanilla_floor_threshold, MultinomialResampler(), term=seq[4])
    assert early.engine.T == 5 and all(p.iteration == 4 for p in early[:3])
