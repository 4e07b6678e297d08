"""oracle/cpu_port (the compiled CPU baseline of bench.py: the hot loop in C++/OpenMP around the per-replica code built by g++)
against the numpy oracle anneal on the same particles, noise and uniforms."""
import numpy as np

from coddiwomple_b200 import testsystems as ts
from coddiwomple_b200.system import spec_from_xml
from oracle import cpu_port, smc as oracle_smc
from oracle.potential import XmlSystemOracle
from tests.helpers import cache_frames, system_xml


def test_cpu_port_tracks_the_oracle_anneal():
    xml = system_xml("fluorobenzene")
    frames = cache_frames()
    x0 = frames[np.random.RandomState(1).randint(0, len(frames), 300)]
    seq = ts.relative_alchemical_sequence(10)
    port = cpu_port.CpuPort(spec_from_xml(xml))
    for thr in (0.5, 0.9999):
        out = port.anneal(x0, seq, seed=5, resample_seed=6, floor_threshold=thr)
        ref = oracle_smc.anneal(XmlSystemOracle(xml), x0, seq, seed=5, resample_seed=6, floor_threshold=thr)
        assert out["resampled_at"] == ref["resampled_at"]
        np.testing.assert_allclose(out["nESS"][0], ref["nESS_trace"][0], rtol=1e-9)
        # same Philox counters: the trajectories agree until the float-cdf search (port) and the integer cdf (oracle) pick a different
        # ancestor at a boundary, i.e. practically always; the estimates agree far inside their statistical error
        assert abs(out["free_energy"] - ref["free_energy"]) < (1e-6 if thr == 0.5 else 0.05)
        assert abs(out["free_energy"] - 1.33) < 0.2                      # the reference's band (tests/legacy_tests.py:90-94)
    assert cpu_port.threads() >= 1
