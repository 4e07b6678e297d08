"""Distribution Factory module — `DistributionFactory`, `TargetFactory`, `ProposalFactory` with the reference's
interface (coddiwomple/distribution_factories.py:16-252)."""
import logging

_logger = logging.getLogger("distribution_factories")


class DistributionFactory():
    """Manage a PDFState."""

    def __init__(self, pdf_state, parameter_sequence, **kwargs):
        self.pdf_state = pdf_state
        self.parameter_sequence = parameter_sequence

    def update_pdf(self, sequence_index):
        parameters = self.parameter_sequence[sequence_index]
        self.pdf_state.set_parameters(parameters)

    def update_parameter_sequence(self, new_parameters):
        new_parameters_type, parameter_sequence_type = type(new_parameters), type(self.parameter_sequence[0])
        assert new_parameters_type == parameter_sequence_type, f"the new_parameters type ({new_parameters_type}) is not the same as the parameter sequence element type ({parameter_sequence_type})"
        self.parameter_sequence.append(new_parameters)


class TargetFactory(DistributionFactory):
    """Manage Target probability distributions."""

    def __init__(self, pdf_state, parameter_sequence, termination_parameters=None, **kwargs):
        super().__init__(pdf_state, parameter_sequence)
        if termination_parameters is None:
            self.termination_parameters = self.parameter_sequence[-1]
        else:
            termination_parameters_type, parameter_type = type(termination_parameters), type(self.parameter_sequence[0])
            assert termination_parameters_type == parameter_type, f"The termination parameters type ({termination_parameters_type}) do not match the sequence parameters type ({parameter_type})"
            self.termination_parameters = termination_parameters

    def compute_incremental_work(self, particle, neglect_proposal_work=False, **kwargs):
        """work_(inc,t) = u_t(x) + auxiliary_work (+ last proposal work).  The flag is inverted in the reference
        (distribution_factories.py:128: `neglect_proposal_work=True` ADDS the proposal work); reproduced deliberately."""
        iteration = particle.iteration
        self.update_pdf(iteration)
        u_t = self.pdf_state.reduced_potential(particle.state)
        state_work = u_t + particle.auxiliary_work
        proposal_work = particle.get_last_proposal_work() if neglect_proposal_work else 0.
        return state_work + proposal_work

    def terminate(self, particle):
        current_particle_parameters = self.parameter_sequence[particle.iteration]
        return True if all(current_particle_parameters[key] == self.termination_parameters[key] for key in current_particle_parameters.keys()) else False


class ProposalFactory(DistributionFactory):
    """Manage Proposal probability distributions"""

    def __init__(self, parameter_sequence, propagator, **kwargs):
        super(ProposalFactory, self).__init__(propagator.pdf_state, parameter_sequence)
        self._propagator = propagator

    def equip_initial_sample(self, particle, initial_particle_state, generation_pdf=None, **kwargs):
        """1. update particle state 2. update proposal work 3. auxiliary_work = -u_0(x_0); returns -log(weight_0)."""
        assert particle.state is None, f"the particle state is not None; it looks like the particle has already been Initialized"
        if generation_pdf is None:
            self.update_pdf(0)
            generation_pdf = self.pdf_state
        # (the reference's else-branch references an undefined name, distribution_factories.py:206; the argument is honoured here)
        state_reduced_potential = self.pdf_state.reduced_potential(initial_particle_state)
        generation_reduced_potential = generation_pdf.reduced_potential(initial_particle_state)
        initial_work = state_reduced_potential - generation_reduced_potential
        particle.update_state(initial_particle_state)
        particle.update_proposal_work(0.)
        particle.update_auxiliary_work(-state_reduced_potential)
        return initial_work

    def propagate(self, particle, num_applications=1, **kwargs):
        iteration = particle.iteration
        self.update_pdf(iteration)
        proposal_work = 0.
        particle_state = particle.state
        for i in range(num_applications):
            particle_state, work = self._propagator.apply(particle_state, **kwargs)
            proposal_work += work
        particle.update_state(particle_state)
        particle.update_proposal_work(proposal_work)
        return particle, proposal_work
