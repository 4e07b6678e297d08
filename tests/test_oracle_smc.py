"""The oracle's whole anneal against the reference's statistical anchors.  CPU only."""
import numpy as np

from coddiwomple_b200 import testsystems as ts
from oracle import langevin, smc
from oracle.potential import XmlSystemOracle, kT_kj_per_mol
from tests.helpers import cache_frames, system_xml

kT = kT_kj_per_mol()


def test_G6_harmonic_oscillator_free_energy_is_one_kT():
    """x0: 0 -> 2 sigma, U0: 0 -> 1 kT  =>  Delta F = 1 kT exactly (coddiwomple/tests/utils.py:63-67)."""
    prm = ts.harmonic_oscillator_parameters()
    o = XmlSystemOracle(ts.harmonic_oscillator_xml())
    r = np.random.RandomState(0)
    x0 = r.normal(0, prm["sigma"], (1000, 1, 3)).astype(np.float32)
    for resampler in ("multinomial", "systematic"):
        out = smc.anneal(o, x0, ts.harmonic_parameter_sequence(21), dt=prm["timestep"], gamma=prm["collision_rate"], n_steps=5, seed=1, resample_seed=2,
                         resampler=resampler)
        assert abs(out["free_energy"] - 1.0) < 0.12, out["free_energy"]


def test_G6_harmonic_oscillator_equilibrium_potential():
    """<u> = 3/2 within 6 sigma_u of theory after 100 x 20 BAOAB steps with velocity re-randomisation
    (coddiwomple/tests/test_openmm.py:192-221)."""
    prm = ts.harmonic_oscillator_parameters()
    o = XmlSystemOracle(ts.harmonic_oscillator_xml())
    x = np.zeros((64, 1, 3))
    v = np.zeros_like(x)
    us = []
    for app in range(100):
        out = langevin.baoab(lambda xx: o.energy_forces(xx), x, v, o.masses, [], prm["timestep"], prm["collision_rate"], kT, 20, seed=5, gids=np.arange(64),
                             step0=20 * app, randomize_velocities=True, track_works=True)
        x, v = out["x"], out["v"]
        us.append(out["u_last"] / kT)
        assert np.all(out["shadow_work"] != 0.0) and np.all(out["proposal_work"] != 0.0)
    us = np.array(us)[10:]
    assert abs(us.mean() - 1.5) < 0.1 and abs(us.std() - np.sqrt(1.5)) < 0.15


def test_G5_fluorobenzene_band():
    """benzene -> fluorobenzene in vacuum, 10 lambdas, 1 BAOAB step per lambda: 1.33 +- 0.2 kT
    (coddiwomple/tests/legacy_tests.py:52-95; troubleshooting/openmm_smc.ipynb JSON 515 logs 1.3376)."""
    o = XmlSystemOracle(system_xml("fluorobenzene"))
    frames = cache_frames()
    r = np.random.RandomState(3)
    x0 = frames[r.randint(0, len(frames), 400)]
    out = smc.anneal(o, x0, ts.relative_alchemical_sequence(10), seed=11, resample_seed=12)
    assert abs(out["free_energy"] - 1.33) < 0.2, out["free_energy"]
    assert np.all(out["labels"] >= 0) and len(out["nESS_trace"]) == 9
