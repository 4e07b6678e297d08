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
    """The library's own resampling kernels when the resampler is one of the stock classes with its stock `_resample`; None for any other
    `Resampler` (base-class SIS included): its `_resample(cumulative_works, num_resamples)` is then called on the host every iteration."""
    for cls, kind in ((SystematicResampler, "systematic"), (MultinomialResampler, "multinomial")):
        if isinstance(resampler, cls) and type(resampler)._resample is cls._resample:
            return kind
    if not isinstance(resampler, Resampler) and not hasattr(resampler, "_resample"):
        raise TypeError("resampler must provide _resample(cumulative_works, num_resamples) (coddiwomple/resamplers.py:202-216)")
    return None


class _WorkView:
    """What an observable reads of a particle (coddiwomple/utils.py:41-64 uses `particle.cumulative_work`)."""
    __slots__ = ("cumulative_work",)

    def __init__(self, w):
        self.cumulative_work = w


def _truncate_at_termination(target_factory):
    """TargetFactory.terminate (distribution_factories.py:134-148) ends the loop at the first member of the sequence that equals
    `termination_parameters`; the stock choice is the last member."""
    seq = list(target_factory.parameter_sequence)
    term = getattr(target_factory, "termination_parameters", None)
    if term is None:
        return seq
    for t, now in enumerate(seq):
        if t > 0 and all(now[k] == term[k] for k in now):
            return seq[:t + 1]
    return seq


def SMC(target_factory, proposal_factory, num_initial_particles, reporter, resampler, initial_positions=None, observable=nESS,
        threshold=vanilla_floor_threshold, floor_threshold=0.5, seed=None, state_factory=None, record_history=False, record_histories=None,
        **kwargs):
    """
    arguments (as the reference)
        target_factory : TargetFactory        — pdf_state + parameter_sequence of the targets (+ termination_parameters)
        proposal_factory : ProposalFactory    — propagator (OMMBIP) + parameter_sequence of the proposals
        num_initial_particles : int
        reporter : Reporter                   — record() is called once at the end (hot path keeps particles on the GPU)
        resampler : Resampler                 — MultinomialResampler / SystematicResampler run on the device; any other Resampler's
                                                `_resample` is called on the host with the works (the particles never leave the GPU)
    extra keyword arguments
        initial_positions : (num_initial_particles, A, 3) array in nm — samples of the first target
        observable / threshold : the stock nESS / vanilla_floor_threshold pair (decision on the device, no host synchronisation), None to
            always resample, or any callables with the reference's signatures `observable(particles, incremental_works, **kwargs)` /
            `threshold(value, **kwargs)` (utils.py:41-72) — these are evaluated on the host once per iteration from the works
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
    stock_rule = always or (observable is nESS and threshold is vanilla_floor_threshold)
    kind = _resampler_kind(resampler)
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1))
    sequence = _truncate_at_termination(target_factory)
    if record_histories is None:
        record_histories = num_initial_particles * len(sequence) <= 50_000_000
    on_device = stock_rule and kind is not None
    engine = SMCEngine(pdf_state.device_system, num_initial_particles, sequence, langevin=propagator.integrator,
                       resampler=kind or "multinomial", floor_threshold=floor_threshold, always_resample=always, seed=seed,
                       resample_seed=seed + 1, record_history=record_history, record_lineage=record_histories, lookahead=on_device)
    if on_device:
        engine.anneal(initial_positions)
    else:
        def decide(cum, inc):
            if always:
                return True
            views = [_WorkView(w) for w in cum]
            return threshold(observable(views, inc, **kwargs), **kwargs)

        draw = None if kind is not None else (lambda W: resampler._resample(cumulative_works=W, num_resamples=len(W), **kwargs))
        engine.initialize(initial_positions)
        for t in range(1, engine.T):
            engine.step_with_host_hooks(t, decide=decide, draw_ancestors=draw)
        engine.finish()
    resampler._resampling_logger.extend(engine.resampled_iterations())
    particles = ParticleList(engine.ensemble, state_factory=state_factory, histories=engine.lineage if record_histories else None)
    particles.engine = engine
    if reporter is not None:
        reporter.record(particles, save_to_disk=True)
    return particles
