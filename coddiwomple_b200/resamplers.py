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
        if isinstance(particles, ParticleList):
            return self._resample_device(particles, incremental_works, observable, threshold, **kwargs)
        if observable is None or threshold is None:
            resample_bool = True
        else:
            _observable_value = observable(particles, incremental_works, **kwargs)
            resample_bool = threshold(_observable_value, **kwargs)
        if resample_bool:
            previous_cumulative_works = np.array([particle.cumulative_work for particle in particles])
            updated_cumulative_works = previous_cumulative_works + np.asarray(incremental_works, dtype=np.float64)
            _, stats, _ = kernels.weight_stats(updated_cumulative_works, always=True)
            mean_cumulative_work = float(stats[_lib.STAT_MEAN].item())  # -logsumexp(-W) + log N (resamplers.py:107)
            resampled_indices = self._resample(cumulative_works=updated_cumulative_works, num_resamples=len(updated_cumulative_works), **kwargs)
            copy_particle_auxiliaries = [particle.auxiliary_work for particle in particles]
            copy_particle_states = [particle.state for particle in particles]
            copy_particle_indices = [particle.ancestry[-1] for particle in particles]
            iterations = [particle.iteration for particle in particles]
            assert all(_iter == iterations[0] for _iter in iterations), f"the particles are desynchronized"
            self._resampling_logger.append(iterations[0])
            for current_particle_index, resampling_particle_index in enumerate(resampled_indices):
                self._update_particle(particles[current_particle_index],
                                      from_particle_auxiliary_work=copy_particle_auxiliaries[resampling_particle_index],
                                      from_particle_state=copy_particle_states[resampling_particle_index],
                                      from_particle_index=copy_particle_indices[resampling_particle_index],
                                      cumulative_work=mean_cumulative_work)
        else:
            [self._null_update_particle(particle, incremental_work, **kwargs) for particle, incremental_work in zip(particles, incremental_works)]

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
        particle.update_state(particle.state)
        particle.update_work(incremental_work)
        particle.update_ancestry(particle.ancestry[-1])

    def _update_particle(self, particle_to_update, from_particle_auxiliary_work, from_particle_state, from_particle_index, cumulative_work,
                         amend_state_history=True, **kwargs):
        particle_to_update.zero_auxiliary_work()
        particle_to_update.update_auxiliary_work(from_particle_auxiliary_work)
        particle_to_update.update_state(copy.deepcopy(from_particle_state), amend_state_history=amend_state_history)
        modified_incremental_work = cumulative_work - particle_to_update.cumulative_work
        particle_to_update.update_work(modified_incremental_work)
        particle_to_update.update_ancestry(from_particle_index)

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
