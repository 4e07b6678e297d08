"""Particle module — `Particle` with the reference's fields and mutators (coddiwomple/particles.py:11-177), plus the
device-backed `ParticleList` the B200 drivers return.

A stand-alone `Particle(index)` is the same host-side bookkeeping object as in the reference (there is no arithmetic
in it to accelerate).  `ParticleList` presents a `ParticleEnsemble` (struct-of-arrays on the GPU) as a sequence of
lazy `Particle` views so that code written against the reference — `[p.cumulative_work for p in particles]`,
`particle.iteration`, `particle.ancestry[-1]` — keeps working without one Python object per particle being built
unless it is asked for.
"""
import copy
import operator

import numpy as np


class Particle:
    """Host-side record of one particle: lineage, works and (optionally) a trace of its states.

    Public surface = the reference's (`coddiwomple/particles.py:11-177`): the `update_*` / `zero_*` mutators and the
    read-only properties generated below.  Storage is a handful of slots; there is nothing here to accelerate.
    """

    __slots__ = ("_w", "_dw", "_pw", "_x", "_line", "_trace", "_aux", "_t")

    def __init__(self, index, record_states=False, iteration=0, **kwargs):
        self._w, self._aux, self._t = 0.0, 0.0, iteration
        self._dw, self._pw = [], []
        self._x = None
        self._line = [index]
        self._trace = [] if record_states else None   # None: states are not recorded

    # ---- works
    def update_work(self, incremental_work):
        self._dw.append(incremental_work)
        self._w += incremental_work

    def update_auxiliary_work(self, incremental_work):
        self._aux += incremental_work

    def zero_auxiliary_work(self):
        self._aux = 0.0

    def update_proposal_work(self, proposal_work):
        self._pw.append(proposal_work)

    def update_proposal_work_in_place(self, proposal_work):
        self._pw[-1] = proposal_work

    def get_last_proposal_work(self):
        return self._pw[-1]

    # ---- lineage / state / clock
    def update_ancestry(self, index):
        self._line.append(index)

    def update_state(self, state, amend_state_history=False, state_history_index=-1):
        if self._trace is not None:
            snapshot = copy.deepcopy(state)
            if amend_state_history:
                self._trace[state_history_index] = snapshot
            else:
                self._trace.append(snapshot)
        self._x = state

    def update_iteration(self, increment=1):
        self._t += increment


# read-only views under the reference's property names
for _name, _slot in (("cumulative_work", "_w"), ("incremental_works", "_dw"), ("proposal_works", "_pw"), ("state", "_x"), ("ancestry", "_line"),
                     ("auxiliary_work", "_aux"), ("state_history", "_trace"), ("iteration", "_t")):
    setattr(Particle, _name, property(operator.attrgetter(_slot)))
del _name, _slot


class _NoHistory(list):
    """Placeholder for a per-iteration history that was not spilled from the device: reading it raises instead of handing back
    a silently truncated list (the reference's lists grow by one entry per iteration, particles.py:38-47)."""

    def __init__(self, what):
        super().__init__()
        self._what = what

    def _fail(self, *a, **k):
        raise RuntimeError(f"Particle.{self._what} was not recorded for this run: pass record_histories=True to SMC() / mcmc_smc_resampler() "
                           "(one (work, label) pair per particle and iteration is then kept on the device and copied out once)")

    __getitem__ = __iter__ = __len__ = __repr__ = __eq__ = _fail


class ParticleList:
    """Sequence view of a device-resident ParticleEnsemble.  Host copies are fetched once, on first access."""

    def __init__(self, ensemble, state_factory=None, histories=None):
        self.ensemble = ensemble
        self._state_factory = state_factory
        self._histories = histories  # optional dict(cum=[T-1][P], labels=[T-1][P]) host arrays, or a callable returning it (fetched lazily)
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
        p._w = float(h["cum"][i])
        p._aux = float(h["aux"][i])
        if callable(self._histories):
            self._histories = self._histories()
        if self._histories is not None:
            cums = self._histories["cum"][:, i]
            p._dw = list(np.diff(np.concatenate([[0.0], cums])))
            p._line = [int(self.ensemble.gid0 + i)] + [int(l) for l in self._histories["labels"][:, i]]
        else:
            p._dw = _NoHistory("incremental_works")
            p._line = _NoHistory("ancestry")
        p._pw = [0.] * (self.ensemble.iteration + 1)
        if self._state_factory is not None:
            p._x = self._state_factory(h["pos"][i], h["vel"][i])
        return p

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]
