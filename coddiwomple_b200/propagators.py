"""Propagator module — abstract interface with the reference's name (coddiwomple/propagators.py:14-39)."""


class Propagator():
    """Generalized propagator object."""

    def __init__(self, **kwargs):
        pass

    def apply(self, particle_state, **kwargs):
        """returns (particle_state, proposal_work); particle_state is modified in place"""
        pass
