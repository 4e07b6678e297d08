"""Host-object API that needs no GPU: `Particle` defaults and mutators (coddiwomple/tests/test_openmm.py:61-92)."""
from coddiwomple_b200.distribution_factories import TargetFactory
from coddiwomple_b200.particles import Particle
from coddiwomple_b200.utils import vanilla_floor_threshold


def test_particle_defaults_and_updates():
    particle = Particle(index=0, record_state=False, iteration=0)  # the reference test passes this typo'd kwarg too
    assert particle.state is None and particle.iteration == 0 and particle.ancestry == [0]
    assert particle.cumulative_work == 0. and particle.auxiliary_work == 0. and particle.incremental_works == []
    particle.update_iteration()
    particle.update_ancestry(1)
    particle.update_work(0.5)
    particle.update_proposal_work(0.)
    particle.update_proposal_work_in_place(0.25)
    particle.update_auxiliary_work(-2.0)
    assert particle.iteration == 1 and particle.ancestry == [0, 1] and particle.cumulative_work == 0.5
    assert particle.get_last_proposal_work() == 0.25 and particle.auxiliary_work == -2.0
    particle.zero_auxiliary_work()
    assert particle.auxiliary_work == 0.
    recording = Particle(index=3, record_states=True)
    recording.update_state({"x": 1})   # the reference crashes here (missing `import copy`, particles.py:138)
    recording.update_state({"x": 2}, amend_state_history=True)
    assert recording.state_history == [{"x": 2}]


def test_threshold_and_termination():
    assert vanilla_floor_threshold(0.5) is True and vanilla_floor_threshold(0.50001) is False
    seq = [{"a": 0.0}, {"a": 0.5}, {"a": 1.0}]
    tf = TargetFactory(pdf_state=None, parameter_sequence=seq)
    p = Particle(0)
    assert not tf.terminate(p)
    p.update_iteration(2)
    assert tf.terminate(p)
