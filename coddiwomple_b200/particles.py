"""Particle module — `Particle` with the reference's fields and mutators (coddiwomple/particles.py:11-177), plus the
device-backed `ParticleList` the B200 drivers return.

A stand-alone `Particle(index)` is the same host-side bookkeeping object as in the reference (there is no arithmetic
in it to accelerate).  `ParticleList` presents a `ParticleEnsemble` (struct-of-arrays on the GPU) as a sequence of
lazy `Particle` views so that code written against the reference — `[p.cumulative_work for p in particles]`,
`particle.iteration`, `particle.ancestry[-1]` — keeps working without one Python object per particle being built
unless it is asked for.
"""
import copy

import numpy as np


class Particle(object):
    """Holds ancestry, proposal works, state, incremental and cumulative works of one particle."""

    def __init__(self, index, record_states=False, iteration=0, **kwargs):
        self._cumulative_work = 0.
        self._incremental_works = []
        self._proposal_works = []
        self._state = None
        self._ancestry = [index]
        self._record_states = record_states
        self._state_history = [] if record_states else None
        self._auxiliary_work = 0.
        self._iteration = iteration

    def update_auxiliary_work(self, incremental_work):
        self._auxiliary_work += incremental_work

    def zero_auxiliary_work(self):
        self._auxiliary_work = 0.

    def update_work(self, incremental_work):
        self._cumulative_work += incremental_work
        self._incremental_works.append(incremental_work)

    def get_last_proposal_work(self):
        return self._proposal_works[-1]

    def update_proposal_work(self, proposal_work):
        self._proposal_works.append(proposal_work)

    def update_proposal_work_in_place(self, proposal_work):
        self._proposal_works[-1] = proposal_work

    def update_ancestry(self, index):
        self._ancestry.append(index)

    def update_state(self, state, amend_state_history=False, state_history_index=-1):
        # the reference calls copy.deepcopy here without importing copy (particles.py:138,140); imported above
        if self._record_states:
            if amend_state_history:
                self._state_history[state_history_index] = copy.deepcopy(state)
            else:
                self._state_history.append(copy.deepcopy(state))
        self._state = state

    def update_iteration(self, increment=1):
        self._iteration += increment

    @property
    def cumulative_work(self):
        return self._cumulative_work

    @property
    def incremental_works(self):
        return self._incremental_works

    @property
    def proposal_works(self):
        return self._proposal_works

    @property
    def state(self):
        return self._state

    @property
    def ancestry(self):
        return self._ancestry

    @property
    def auxiliary_work(self):
        return self._auxiliary_work

    @property
    def state_history(self):
        return self._state_history

    @property
    def iteration(self):
        return self._iteration


class ParticleList:
    """Sequence view of a device-resident ParticleEnsemble.  Host copies are fetched once, on first access."""

    def __init__(self, ensemble, state_factory=None, histories=None):
        self.ensemble = ensemble
        self._state_factory = state_factory
        self._histories = histories  # optional dict(cum=[T][P], labels=[T][P]) host arrays
        self._host = None

    def __len__(self):
        return self.ensemble.P

    def refresh(self):
        self._host = None

    def _fetch(self):
        if self._host is None:
            e = self.ensemble
            self._host = dict(cum=e.cum.cpu().numpy(), aux=e.auxiliary_work.cpu().numpy(), label=e.labels.cpu().numpy(),
                              pos=e.positions.cpu().numpy(), vel=e.velocities.cpu().numpy())
        return self._host

    @property
    def cumulative_works(self):
        return self._fetch()["cum"]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        h = self._fetch()
        if i < 0:
            i += len(self)
        p = Particle(index=int(self.ensemble.gid0 + i), iteration=self.ensemble.iteration)
        p._cumulative_work = float(h["cum"][i])
        p._auxiliary_work = float(h["aux"][i])
        if self._histories is not None:
            cums = self._histories["cum"][:, i]
            p._incremental_works = list(np.diff(np.concatenate([[0.0], cums])))
            p._ancestry = [int(self.ensemble.gid0 + i)] + [int(l) for l in self._histories["labels"][:, i]]
        else:
            p._ancestry = [int(self.ensemble.gid0 + i), int(h["label"][i])]
        p._proposal_works = [0.] * (self.ensemble.iteration + 1)
        if self._state_factory is not None:
            p._state = self._state_factory(h["pos"][i], h["vel"][i])
        return p

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]
