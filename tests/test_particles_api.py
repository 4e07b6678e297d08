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


def test_lepton_protocol_functions():
    """The alchemical functions of the reference's AIS test (coddiwomple/tests/utils.py:100-110) evaluated by the host-side
    Lepton evaluator equal the piecewise-linear protocol of its SMC test (coddiwomple/tests/legacy_tests.py:58-76)."""
    import numpy as np
    import pytest
    from coddiwomple_b200.openmm.lepton import LeptonError, compile_expression, protocol_table
    from coddiwomple_b200 import testsystems as ts
    x = 'fractional_iteration'
    alchemical_functions = {'lambda_sterics_core': x, 'lambda_electrostatics_core': x,
                            'lambda_sterics_insert': f"select(step({x} - 0.5), 1.0, 2.0 * {x})",
                            'lambda_sterics_delete': f"select(step({x} - 0.5), 2.0 * ({x} - 0.5), 0.0)",
                            'lambda_electrostatics_insert': f"select(step({x} - 0.5), 2.0 * ({x} - 0.5), 0.0)",
                            'lambda_electrostatics_delete': f"select(step({x} - 0.5), 1.0, 2.0 * {x})",
                            'lambda_bonds': x, 'lambda_angles': x, 'lambda_torsions': x}
    table = protocol_table(alchemical_functions, 8)
    want = ts.relative_alchemical_sequence(9)[1:]
    for got, ref in zip(table, want):
        for k in ref:
            assert got[k] == pytest.approx(ref[k], abs=1e-15)
    assert compile_expression("-2^2 + 3*(4-1)/2 - -1")() == 1.5
    assert compile_expression("(1-f)*0.0 + f*0.2")(f=0.5) == pytest.approx(0.1)
    with pytest.raises(LeptonError):
        compile_expression("foo(1)")
    with pytest.raises(LeptonError):
        compile_expression("1 +")
    harmonic = {f"{ts.HO_PREFIX}_x0": f"(1-{x})*0.0 + {x}*0.2", f"{ts.HO_PREFIX}_U0": f"(1-{x})*0.0 + {x}*2.4943386436406856"}
    rows = protocol_table(harmonic, 20)
    assert rows[-1][f"{ts.HO_PREFIX}_x0"] == pytest.approx(0.2) and np.isclose(rows[9][f"{ts.HO_PREFIX}_U0"], 0.5 * 2.4943386436406856)


def test_reporter_writes_the_reference_wire_format(tmp_path):
    """OpenMMReporter without mdtraj: one multi-model PDB per particle, readable by the cache reader (round trip at the
    3-decimal Angstrom resolution of the format); no directory -> no-op (coddiwomple/tests/legacy_tests.py:83-84)."""
    import numpy as np
    from coddiwomple_b200.openmm.reporters import OpenMMReporter
    from coddiwomple_b200.openmm.utils import read_pdb_frames

    class S:
        def __init__(self, x):
            self.positions = x

    rng = np.random.RandomState(0)
    particles = [Particle(i) for i in range(3)]
    reporter = OpenMMReporter(str(tmp_path), "traj", atom_names=["C1", "C2", "H1", "F1"])
    frames = []
    for step in range(4):
        xs = rng.normal(0, 0.2, (3, 4, 3))
        frames.append(xs)
        for p, x in zip(particles, xs):
            p.update_state(S(x))
        reporter.record(particles, save_to_disk=(step == 3))
    assert reporter.hex_counter == len(reporter.hex_dict) == 3
    for i in range(3):
        back = read_pdb_frames(str(tmp_path / f"traj.{i:04d}.pdb"))
        assert back.shape == (4, 4, 3)
        np.testing.assert_allclose(back, np.stack(frames)[:, i], atol=6e-5)
    OpenMMReporter(None, None).record(particles, save_to_disk=True)  # no-op
