"""SMC module — `SMC(target_factory, proposal_factory, num_initial_particles, reporter, resampler, **kwargs)`.

In the reference this function is a docstring with no body (coddiwomple/smc.py:16-69); the loop it sketches is the
one `mcmc_smc_resampler` executes (coddiwomple/openmm/coddiwomple.py:312-331).  Here it has a body: the loop runs on
the GPU through `SMCEngine` (struct-of-arrays particles, fused kernels, on-device resampling decision).
"""
import numpy as np

from .engine import SMCEngine
from .particles import ParticleList
from .resamplers import MultinomialResampler, Resampler, SystematicResampler
from .utils import nESS, vanilla_floor_threshold


def _resampler_kind(resampler):
    if isinstance(resampler, SystematicResampler):
        return "systematic"
    if isinstance(resampler, MultinomialResampler):
        return "multinomial"
    raise NotImplementedError("the device loop supports MultinomialResampler and SystematicResampler")


def SMC(target_factory, proposal_factory, num_initial_particles, reporter, resampler, initial_positions=None, observable=nESS,
        threshold=vanilla_floor_threshold, floor_threshold=0.5, seed=None, state_factory=None, record_history=False, record_histories=None,
        **kwargs):
    """
    arguments (as the reference)
        target_factory : TargetFactory        — pdf_state + parameter_sequence of the targets
        proposal_factory : ProposalFactory    — propagator (OMMBIP) + parameter_sequence of the proposals
        num_initial_particles : int
        reporter : Reporter                   — record() is called once at the end (hot path keeps particles on the GPU)
        resampler : MultinomialResampler | SystematicResampler
    extra keyword arguments
        initial_positions : (num_initial_particles, A, 3) array in nm — samples of the first target
        observable / threshold : the stock nESS / vanilla_floor_threshold pair, or None to always resample
        record_histories : keep every particle's per-iteration cumulative work and root label on the device so that
            `particle.incremental_works` / `particle.ancestry` have one entry per iteration like the reference's
            (particles.py:38-47).  None (default): on when particles x iterations <= 5e7 (12 bytes each), else off — reading those
            two properties then raises instead of returning truncated lists.

    returns
        particles : ParticleList (sequence of Particle views of the device-resident ensemble)
    """
    if initial_positions is None:
        raise ValueError("initial_positions (samples of the first distribution) are required")
    pdf_state = target_factory.pdf_state
    propagator = proposal_factory._propagator
    always = observable is None or threshold is None
    if not always and not (observable is nESS and threshold is vanilla_floor_threshold):
        raise NotImplementedError("the device loop supports the stock nESS / vanilla_floor_threshold pair (or None)")
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1))
    if record_histories is None:
        record_histories = num_initial_particles * len(target_factory.parameter_sequence) <= 50_000_000
    engine = SMCEngine(pdf_state.device_system, num_initial_particles, target_factory.parameter_sequence, langevin=propagator.integrator,
                       resampler=_resampler_kind(resampler), floor_threshold=floor_threshold, always_resample=always, seed=seed,
                       resample_seed=seed + 1, record_history=record_history, record_lineage=record_histories)
    engine.anneal(initial_positions)
    resampler._resampling_logger.extend(engine.resampled_iterations())
    particles = ParticleList(engine.ensemble, state_factory=state_factory, histories=engine.lineage if record_histories else None)
    particles.engine = engine
    if reporter is not None:
        reporter.record(particles, save_to_disk=True)
    return particles
