"""Resampler module — `Resampler`, `MultinomialResampler` (and the new `SystematicResampler`) with the reference's
interface (coddiwomple/resamplers.py:16-272), executed by the K2/K3 CUDA kernels.

`resample(particles, incremental_works, observable, threshold)`:
  * device path — `particles` is a `ParticleList` and the observable/threshold are the stock `nESS` /
    `vanilla_floor_threshold` (or None): everything (log-sum-exp, nESS, decision, integer cdf, search, cum update)
    runs on the GPU without a host round trip; the state copy is deferred to the next gather-on-load;
  * host-object path — `particles` is a list of `Particle`: works are staged to the GPU, ancestors come from the same
    kernels, and the per-particle bookkeeping of `_update_particle` / `_null_update_particle` is applied to the
    Python objects exactly as the reference does.
"""
import copy
import logging

import numpy as np

from . import _lib, kernels
from .particles import ParticleList
from .utils import nESS, vanilla_floor_threshold

_logger = logging.getLogger("resamplers")


class Resampler():
    """Base class: `_resample` does NOT resample (sequential importance sampling), as in the reference."""
    kind = None  # identity

    def __init__(self, seed=None, **kwargs):
        self._resampling_logger = list()
        self._seed = seed
        self._draws = 0

    @property
    def resampling_logger(self):
        return copy.deepcopy(self._resampling_logger)

    # ------------------------------------------------------------------ uniforms
    def _next_uniforms(self, n):
        """Uniforms for one draw.  The reference consumes the GLOBAL numpy stream (np.random.choice,
        resamplers.py:271); so does this, unless a seed was given, in which case they come from Philox on the device."""
        self._draws += 1
        if self._seed is None:
            return np.random.random_sample(n)
        return kernels.philox_uniforms(self._seed, self._draws, n)

    # ------------------------------------------------------------------ public API
    def resample(self, particles, incremental_works, observable=None, threshold=None, update_particle_indices=True, **kwargs):
        """Decide (observable + threshold, or always when either is None) and, if so, resample; otherwise every particle
        just accumulates its incremental work (coddiwomple/resamplers.py:61-133)."""
        if isinstance(particles, ParticleList):
            return self._resample_device(particles, incremental_works, observable, threshold, **kwargs)
        wanted = True
        if observable is not None and threshold is not None:
            wanted = threshold(observable(particles, incremental_works, **kwargs), **kwargs)
        if not wanted:
            for particle, work in zip(particles, incremental_works):
                self._null_update_particle(particle, work, **kwargs)
            return
        self._host_resample(particles, np.asarray(incremental_works, dtype=np.float64), **kwargs)

    def _host_resample(self, particles, incremental_works, **kwargs):
        """Python `Particle` objects: the works go to the GPU, ancestors and the normaliser -logsumexp(-W) + log N
        (resamplers.py:107) come back from the K2/K3 kernels, and each particle is rewritten from a snapshot of its ancestor."""
        clock = {p.iteration for p in particles}
        if len(clock) != 1:
            raise AssertionError(f"the particles are desynchronized (iterations {sorted(clock)})")
        works = np.fromiter((p.cumulative_work for p in particles), dtype=np.float64, count=len(particles)) + incremental_works
        _, stats, _ = kernels.weight_stats(works, always=True)
        normaliser = float(stats[_lib.STAT_MEAN].item())
        ancestors = self._resample(cumulative_works=works, num_resamples=len(works), **kwargs)
        snapshot = [(p.auxiliary_work, p.state, p.ancestry[-1]) for p in particles]   # taken before anything is overwritten
        self._resampling_logger.append(clock.pop())
        for particle, a in zip(particles, ancestors):
            aux, state, label = snapshot[int(a)]
            self._update_particle(particle, from_particle_auxiliary_work=aux, from_particle_state=state, from_particle_index=label,
                                  cumulative_work=normaliser)

    def _resample_device(self, particles, incremental_works, observable, threshold, floor_threshold=0.5, **kwargs):
        import torch
        if self.kind is None:
            raise NotImplementedError("the SIS base class has no device path; use a list of Particle objects")
        e = particles.ensemble
        stock = (observable is None or threshold is None) or (observable is nESS and threshold is vanilla_floor_threshold)
        if not stock:
            raise NotImplementedError("device particles support the stock nESS / vanilla_floor_threshold pair (or None); "
                                      "pass a list of Particle objects for custom observables")
        always = observable is None or threshold is None
        inc = incremental_works if torch.is_tensor(incremental_works) else torch.as_tensor(np.asarray(incremental_works, dtype=np.float64)).to(e.device)
        e.wmin_key.fill_(-1)
        _lib.check(e.lib.cdw_weight_update(_lib.ptr(inc), None, None, _lib.ptr(e.cum), e.P, _lib.ptr(e.inc), _lib.ptr(e.W), _lib.ptr(e.wmin_key),
                                          _lib.current_stream()), "cdw_weight_update")
        stats = torch.zeros(_lib.STATS_LEN, dtype=torch.float64, device=e.device)
        e.weight_stats(stats, -1.0 if always else floor_threshold)
        n_u = e.P if self.kind == _lib.CDW_RESAMPLE_MULTINOMIAL else 1
        u = self._next_uniforms(n_u)
        u = u if torch.is_tensor(u) else torch.as_tensor(u).to(e.device)
        e.resample(stats, self.kind, uniforms=u)
        if float(stats[_lib.STAT_RESAMPLE].item()) != 0.0:
            self._resampling_logger.append(e.iteration)
        particles.refresh()
        return stats

    def _null_update_particle(self, particle, incremental_work, **kwargs):
        """No resampling: the particle keeps its state and lineage and accumulates the work (resamplers.py:136-156)."""
        particle.update_work(incremental_work)
        particle.update_state(particle.state)
        particle.update_ancestry(particle.ancestry[-1])

    def _update_particle(self, particle_to_update, from_particle_auxiliary_work, from_particle_state, from_particle_index, cumulative_work,
                         amend_state_history=True, **kwargs):
        """Overwrite a particle with (a copy of) its ancestor; its cumulative work becomes the ensemble normaliser, i.e. the
        incremental work recorded is (normaliser - previous cumulative work) (resamplers.py:159-199)."""
        target = particle_to_update
        target.update_work(cumulative_work - target.cumulative_work)
        target.zero_auxiliary_work()
        target.update_auxiliary_work(from_particle_auxiliary_work)
        target.update_ancestry(from_particle_index)
        target.update_state(copy.deepcopy(from_particle_state), amend_state_history=amend_state_history)

    def _resample(self, cumulative_works, num_resamples, **kwargs):
        """Dummy _resample method; does NOT resample (coddiwomple/resamplers.py:202-216)."""
        return range(num_resamples)


class MultinomialResampler(Resampler):
    """Vanilla multinomial resampling: the outcome of np.random.choice(N, N, p=softmax(-works)) for the same uniforms
    (coddiwomple/resamplers.py:254-272), computed by the integer-cdf scan + search kernels."""
    kind = _lib.CDW_RESAMPLE_MULTINOMIAL

    def _resample(self, cumulative_works, num_resamples, **kwargs):
        if num_resamples != len(cumulative_works):
            raise NotImplementedError("a variable number of resamples is a TODO in the reference as well")
        u = self._next_uniforms(num_resamples)
        return kernels.resample_indices(np.asarray(cumulative_works, dtype=np.float64), u, kind="multinomial")


class SystematicResampler(Resampler):
    """Systematic resampling u_i = (u_0 + i)/N — not in the reference; named by BASELINE.json."""
    kind = _lib.CDW_RESAMPLE_SYSTEMATIC

    def _resample(self, cumulative_works, num_resamples, **kwargs):
        if num_resamples != len(cumulative_works):
            raise NotImplementedError("a variable number of resamples is not supported")
        u = self._next_uniforms(1)
        return kernels.resample_indices(np.asarray(cumulative_works, dtype=np.float64), u, kind="systematic")
