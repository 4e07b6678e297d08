"""The oracle's analytic forces against central differences, for every vacuum fixture and the synthetic systems."""
import numpy as np
import pytest

from tests.helpers import VACUUM_SYSTEMS, hybrid_positions, jittered, system_xml
from oracle.potential import XmlSystemOracle


def _fd(o, x, lam, h=1e-6):
    num = np.zeros_like(x)
    for a in range(x.shape[0]):
        for c in range(3):
            xp, xm = x.copy(), x.copy()
            xp[a, c] += h
            xm[a, c] -= h
            num[a, c] = -(o.potential_energy(xp, lam) - o.potential_energy(xm, lam)) / (2 * h)
    return num


@pytest.mark.parametrize("name", VACUUM_SYSTEMS)
def test_forces_match_finite_differences(name):
    o = XmlSystemOracle(system_xml(name))
    x = jittered(hybrid_positions(name), 1, seed=3, scale=0.005)[0]
    rng = np.random.RandomState(4)
    lam = {k: float(rng.uniform(0.05, 0.95)) for k in o.global_defaults if k.startswith("lambda")}
    _, F = o.energy_forces(x, lam)
    assert np.abs(_fd(o, x, lam) - F).max() < 2e-6 * max(1.0, np.abs(F).max())


def test_synthetic_systems_forces():
    from coddiwomple_b200 import testsystems as ts
    xml, x = ts.alanine_dipeptide_shaped_xml()
    o = XmlSystemOracle(xml)
    x = jittered(x, 1, seed=1, scale=0.003)[0]
    lam = ts.relative_alchemical_sequence(7)[4]
    _, F = o.energy_forces(x, lam)
    assert np.abs(_fd(o, x, lam) - F).max() < 2e-6 * np.abs(F).max()
    oh = XmlSystemOracle(ts.harmonic_oscillator_xml())
    xh = np.array([[0.03, -0.02, 0.05]])
    lamh = {f"{ts.HO_PREFIX}_x0": 0.1, f"{ts.HO_PREFIX}_U0": 1.0}
    U, F = oh.energy_forces(xh, lamh)
    K = oh.global_defaults[f"{ts.HO_PREFIX}_K"]
    assert U == pytest.approx(0.5 * K * ((0.03 - 0.1) ** 2 + 0.02 ** 2 + 0.05 ** 2) + 1.0)
    np.testing.assert_allclose(F, -K * (xh - [[0.1, 0, 0]]))


def test_soft_core_cap_is_continuous():
    """U_sterics and dU/dr are continuous at r = r_LJ (the quadratic cap of the perses expression)."""
    sigma, eps, rlj = 0.33, 0.4, 0.29
    r = np.array([rlj - 1e-9, rlj + 1e-9])
    e, d = XmlSystemOracle.softcore_lj(r, sigma, eps, np.array([rlj, rlj]))
    assert abs(e[0] - e[1]) < 1e-6 and abs(d[0] - d[1]) < 1e-4
