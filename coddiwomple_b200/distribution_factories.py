"""Distribution Factory module — `DistributionFactory`, `TargetFactory`, `ProposalFactory` with the reference's
interface (coddiwomple/distribution_factories.py:16-252): a factory owns a `PDFState` and the sequence of parameter sets
it is annealed through; the target factory turns a particle into an incremental work, the proposal factory moves it."""
import logging

_logger = logging.getLogger("distribution_factories")


def _same_kind(candidate, sequence, what):
    have, want = type(candidate), type(sequence[0])
    if have is not want:
        raise AssertionError(f"{what}: got {have.__name__}, the sequence holds {want.__name__}")


class DistributionFactory:
    """A `pdf_state` plus the list of parameter sets that index its members (`parameter_sequence[t]`)."""

    def __init__(self, pdf_state, parameter_sequence, **kwargs):
        self.pdf_state, self.parameter_sequence = pdf_state, parameter_sequence

    def update_pdf(self, sequence_index):
        """Point the pdf_state at member `sequence_index` of the sequence."""
        self.pdf_state.set_parameters(self.parameter_sequence[sequence_index])

    def update_parameter_sequence(self, new_parameters):
        """Append one more member (on-the-fly protocols)."""
        _same_kind(new_parameters, self.parameter_sequence, "update_parameter_sequence")
        self.parameter_sequence.append(new_parameters)


class TargetFactory(DistributionFactory):
    """Target distributions: incremental works and the termination test."""

    def __init__(self, pdf_state, parameter_sequence, termination_parameters=None, **kwargs):
        super().__init__(pdf_state, parameter_sequence)
        if termination_parameters is not None:
            _same_kind(termination_parameters, self.parameter_sequence, "termination_parameters")
        self.termination_parameters = self.parameter_sequence[-1] if termination_parameters is None else termination_parameters

    def compute_incremental_work(self, particle, neglect_proposal_work=False, **kwargs):
        """u_t(x) + auxiliary_work, where auxiliary_work = -u_{t-1}(x_{t-1}) is carried by the particle.

        The flag reads backwards in the reference (distribution_factories.py:128 ADDS the last proposal work when
        `neglect_proposal_work` is true); that behaviour is kept on purpose so that callers see the same numbers."""
        self.update_pdf(particle.iteration)
        work = self.pdf_state.reduced_potential(particle.state) + particle.auxiliary_work
        if neglect_proposal_work:
            work += particle.get_last_proposal_work()
        return work

    def terminate(self, particle):
        """True once the particle's current parameter set equals the termination set, key by key."""
        now = self.parameter_sequence[particle.iteration]
        return all(now[k] == self.termination_parameters[k] for k in now)


class ProposalFactory(DistributionFactory):
    """Proposal distributions: equips the first sample and applies the propagator."""

    def __init__(self, parameter_sequence, propagator, **kwargs):
        super().__init__(propagator.pdf_state, parameter_sequence)
        self._propagator = propagator

    def equip_initial_sample(self, particle, initial_particle_state, generation_pdf=None, **kwargs):
        """Give a fresh particle its state, a zero proposal work and auxiliary_work = -u_0(x_0); returns the initial work
        u_0(x_0) - u_gen(x_0) (zero when the sample was drawn from the first target itself).
        (The reference's branch for an explicit `generation_pdf` uses an undefined name, distribution_factories.py:206;
        here the argument simply is the generating distribution.)"""
        if particle.state is not None:
            raise AssertionError("equip_initial_sample: this particle already has a state")
        self.update_pdf(0)
        u_target = self.pdf_state.reduced_potential(initial_particle_state)
        u_generation = u_target if generation_pdf is None else generation_pdf.reduced_potential(initial_particle_state)
        particle.update_state(initial_particle_state)
        particle.update_proposal_work(0.0)
        particle.update_auxiliary_work(-u_target)
        return u_target - u_generation

    def propagate(self, particle, num_applications=1, **kwargs):
        """Apply the propagator `num_applications` times under the particle's current parameter set; the summed proposal
        work is stored in the particle and returned with it."""
        self.update_pdf(particle.iteration)
        state, total = particle.state, 0.0
        for _ in range(num_applications):
            state, work = self._propagator.apply(state, **kwargs)
            total += work
        particle.update_state(state)
        particle.update_proposal_work(total)
        return particle, total
