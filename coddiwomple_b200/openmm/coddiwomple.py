"""Drivers with the reference's names (coddiwomple/openmm/coddiwomple.py): `mcmc_smc_resampler`, `handle_pressure`.

`mcmc_smc_resampler` keeps the reference's signature and defaults (:238-254) and executes its loop (:312-331) on the
GPU.  Differences forced by the environment: `system` is OpenMM System XML / a SystemSpec (an openmm.System is
serialised when OpenMM is installed); `endstate_cache_filename` may be the PDB cache, a .npy file or an array;
`md_topology`, `directory_name`, `trajectory_prefix` only matter for trajectory writing, which is out of scope.
"""
import numpy as np

from ..distribution_factories import ProposalFactory, TargetFactory
from ..engine import as_spec
from ..resamplers import MultinomialResampler
from ..smc import SMC
from ..utils import nESS, vanilla_floor_threshold
from .integrators import OMMLI, OMMLIAIS
from .propagators import OMMAISPR, OMMBIP
from .reporters import OpenMMReporter
from .states import OpenMMParticleState, OpenMMPDFState, RelativeAlchemicalState
from .utils import load_frames

DEFAULT_INTEGRATOR_KWARGS = {'temperature': 300.0, 'collision_rate': 1.0, 'timestep': 0.002, 'splitting': "V R O R V", 'constraint_tolerance': 1e-6}


def handle_pressure(system):
    """None for vacuum systems, 1 atm if the system has a MonteCarloBarostat (openmm/coddiwomple.py:124-143)."""
    xml = getattr(as_spec(system), "xml", None) or ""
    return 1.0 if 'type="MonteCarloBarostat"' in xml else None


def annealed_importance_sampling(system, endstate_cache_filename, directory_name, trajectory_prefix, md_topology, alchemical_functions,
                                 number_of_applications, steps_per_application, endstate_parameters, alchemical_composability=RelativeAlchemicalState,
                                 integrator_kwargs={'temperature': 300.0, 'collision_rate': 1.0, 'timestep': 0.002, 'splitting': "P I H N S V R O R V",
                                                    'constraint_tolerance': 1e-6},
                                 record_state_work_interval=1, trajectory_write_interval=1, seed=None):
    """Annealed importance sampling (coddiwomple/openmm/coddiwomple.py:146-236): `number_of_applications` independent
    non-equilibrium switching trajectories from random end-state cache frames; returns {trajectory index: [state works in kT]}.
    All trajectories run at once on the GPU (one replica per team of lanes) instead of one after the other."""
    pressure = handle_pressure(system)
    print(f"pressure: {pressure}")
    frames = load_frames(endstate_cache_filename)
    num_frames = frames.shape[0]
    pdf_state = OpenMMPDFState(system=system, alchemical_composability=alchemical_composability, pressure=pressure)
    pdf_state_parameters = pdf_state.get_parameters()
    reset_pdf_state_parameters = {key: endstate_parameters for key in pdf_state_parameters.keys()}
    pdf_state.set_parameters(reset_pdf_state_parameters)
    reporter = OpenMMReporter(directory_name, trajectory_prefix, md_topology=md_topology)
    ais_integrator = OMMLIAIS(alchemical_functions, steps_per_application, **integrator_kwargs)
    ais_propagator = OMMAISPR(openmm_pdf_state=pdf_state, integrator=ais_integrator, record_state_work_interval=record_state_work_interval,
                              reporter=reporter, trajectory_write_interval=trajectory_write_interval, context_cache=None, reassign_velocities=True,
                              n_restart_attempts=0, seed=seed)
    chosen = np.random.choice(range(num_frames), number_of_applications)
    works, _ = ais_propagator.apply_batch(frames[chosen], n_steps=steps_per_application)
    for i in range(number_of_applications):
        ais_propagator._state_works[i] = [float(w) for w in works[i]]
    ais_propagator._state_works_counter = number_of_applications
    return ais_propagator.state_works


def mcmc_smc_resampler(system, endstate_cache_filename, directory_name, trajectory_prefix, md_topology, number_of_particles, parameter_sequence,
                       observable=nESS, threshold=vanilla_floor_threshold, alchemical_composability=RelativeAlchemicalState,
                       integrator_kwargs=DEFAULT_INTEGRATOR_KWARGS, record_state_work_interval=1, trajectory_write_interval=1, resampler=None,
                       seed=None, record_histories=None):
    """Single-direction annealing SMC in the AIS regime with resampling; returns the particles (ParticleList)."""
    pressure = handle_pressure(system)
    print(f"pressure: {pressure}")
    frames = load_frames(endstate_cache_filename)
    num_frames = frames.shape[0]
    proposal_pdf_state = OpenMMPDFState(system=system, alchemical_composability=alchemical_composability, pressure=pressure)
    target_pdf_state = OpenMMPDFState(system=proposal_pdf_state.device_system, alchemical_composability=alchemical_composability, pressure=pressure)
    reporter = OpenMMReporter(directory_name, trajectory_prefix, md_topology=md_topology)
    integrator = OMMLI(**integrator_kwargs)
    propagator = OMMBIP(proposal_pdf_state, integrator)
    proposal_factory = ProposalFactory(parameter_sequence, propagator)
    target_factory = TargetFactory(target_pdf_state, parameter_sequence, termination_parameters=None)
    resampler = resampler if resampler is not None else MultinomialResampler()
    # one random cache frame per particle (openmm/coddiwomple.py:302-304; global numpy stream, as the reference)
    random_frames = np.array([np.random.choice(range(num_frames)) for _ in range(number_of_particles)]) if number_of_particles <= 4096 \
        else np.random.randint(0, num_frames, number_of_particles)
    initial_positions = frames[random_frames]
    particles = SMC(target_factory, proposal_factory, number_of_particles, reporter, resampler, initial_positions=initial_positions,
                    observable=observable, threshold=threshold, seed=seed, record_histories=record_histories,
                    state_factory=lambda x, v: OpenMMParticleState(positions=x, velocities=v))
    return particles
