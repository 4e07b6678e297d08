"""Pin the oracle against what the reference itself recorded.  CPU only."""
import numpy as np
import pytest
from scipy.special import logsumexp

from tests.helpers import generic_layer, reference_numbers, relative_lambdas, system_xml, eq0000_frames
from oracle import resampling, rng, weights
from oracle.potential import XmlSystemOracle, kT_kj_per_mol


@pytest.fixture(scope="module")
def fluoro():
    return XmlSystemOracle(system_xml("fluorobenzene"))


def test_kT_constant():
    # kB*N_A*300 K with the OpenMM 7.4 constants
    assert kT_kj_per_mol(300.0) == pytest.approx(2.4943386436406856, rel=1e-15)


def test_G1_G3_reduced_potentials_logged_by_the_reference(fluoro):
    """troubleshooting/openmm_smc.ipynb JSON 260/264/268: u(all lambda = 0) of frames 2, 0, 1 of tester/eq.0000.pdb.
    The reference ran OpenMM's single-precision CPU platform, so the fp64 restatement sits 1-2e-6 away."""
    ref = reference_numbers()["u_lambda0"]
    u = fluoro.reduced_potential(eq0000_frames()[:3].astype(np.float64))
    for key, frame in (("frame2", 2), ("frame0", 0), ("frame1", 1)):
        assert u[frame] == pytest.approx(ref[key][0], rel=5e-6)


def test_G4_first_iteration_normaliser(fluoro):
    """openmm_smc.ipynb JSON 339: -LSE(-dw) + log 3 for lambda_0 -> lambda_1 (q = 1/9) on those three frames."""
    x = eq0000_frames()[:3].astype(np.float64)
    dw = fluoro.reduced_potential(x, relative_lambdas(1.0 / 9.0)) - fluoro.reduced_potential(x)
    mean = -logsumexp(-dw) + np.log(3)
    assert mean == pytest.approx(reference_numbers()["first_mean_cumulative_work"][0], rel=5e-6)
    # per-frame values recorded by the survey probe
    np.testing.assert_allclose(dw[[2, 0, 1]], [0.25116192, 0.25021867, 0.25282788], atol=2e-7)


def test_block_energies_frame0(fluoro):
    """per-block energies (kJ/mol) at lambda = 0 for frame 0."""
    _, _, parts = fluoro.energy_forces(eq0000_frames()[0].astype(np.float64), per_block=True)
    got = {k: float(e[0]) for k, e in parts}
    want = dict(cbond=0.08439, hbond=0.01171, cangle=0.00952, hangle=1.84271, ptorsion=0.00804, nonbonded=18.52195, csterics=-0.74404)
    for k, v in want.items():
        assert got[k] == pytest.approx(v, abs=6e-6)


@pytest.mark.parametrize("case", range(6))
def test_generic_layer_bit_exact(case):
    """oracle/weights.py + oracle/resampling.py reproduce the reference's importable generic layer bit for bit
    (tests/golden/generic_layer.json was produced by importing coddiwomple.{utils,resamplers,particles})."""
    c = generic_layer()["cases"][case]
    cum, inc, aux = (np.array(c[k]) for k in ("cumulative_works_in", "incremental_works", "auxiliary_works_in"))
    labels = np.arange(c["P"])
    assert weights.nESS(cum, inc) == c["nESS"]
    assert weights.vanilla_floor_threshold(c["nESS"]) == c["threshold_says_resample"]
    np.testing.assert_array_equal(weights.normalized_weights(cum + inc), np.array(c["normalized_weights"]))
    u = np.array(c["uniforms"])
    for always, key in ((False, "after_thresholded"), (True, "after_forced")):
        for kind in ("float", "multinomial"):  # numpy's own procedure, and the integer-cdf form the GPU implements
            out = weights.resample_step(cum, inc, aux, labels, uniforms=u, kind=kind, always=always)
            np.testing.assert_array_equal(out["cum"], np.array(c[key]["cumulative_work"]))
            np.testing.assert_array_equal(out["aux"], np.array(c[key]["auxiliary_work"]))
            np.testing.assert_array_equal(out["labels"], np.array(c[key]["ancestry_last"]))
            np.testing.assert_array_equal(out["ancestors"], np.array(c[key]["state_label"]))


def test_sis_base_class():
    c = generic_layer()["sis_case"]
    inc = np.array(c["incremental_works"])
    out = weights.resample_step(np.zeros(len(inc)), inc, np.zeros(len(inc)), np.arange(len(inc)), kind="identity", always=True)
    np.testing.assert_array_equal(out["cum"], np.array(c["cumulative_work"]))
    np.testing.assert_array_equal(out["labels"], np.array(c["ancestry_last"]))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_G7_np_random_choice_equals_cdf_search(seed):
    """np.random.choice(N, N, p) == searchsorted(cumsum(p)/cumsum(p)[-1], random_sample(N), 'right')
    and the integer-cdf form agrees with both."""
    r = np.random.RandomState(seed)
    W = r.normal(0, 2.0, 5000)
    p = weights.normalized_weights(W)
    np.random.seed(seed + 10)
    ref = np.random.choice(len(W), len(W), p=p, replace=True)
    np.random.seed(seed + 10)
    u = np.random.random_sample(len(W))
    np.testing.assert_array_equal(resampling.multinomial_float(W, u), ref)
    np.testing.assert_array_equal(resampling.ancestors_fixed(W, u), ref)


def test_fixed_point_matches_float_path_large():
    """The count of ancestors that differ between numpy's fp64 cumsum path and the integer cdf: expected 0 up to 1e6."""
    r = np.random.RandomState(5)
    for n, s in ((100_000, 0.1), (100_000, 3.0), (1_000_000, 1.0)):
        W, u = r.normal(0, s, n), r.random_sample(n)
        assert int((resampling.multinomial_float(W, u) != resampling.ancestors_fixed(W, u)).sum()) == 0


def test_det_exp_accuracy_and_edges():
    x = -np.abs(np.random.RandomState(0).normal(0, 40, 100000))
    x[:4] = [0.0, -699.9, -700.1, -1e6]
    e = resampling.det_exp(x)
    ok = x >= -700
    assert np.max(np.abs(e[ok] - np.exp(x[ok])) / np.exp(x[ok])) < 2.5e-16
    assert e[0] == 1.0 and e[2] == 0.0 and e[3] == 0.0


def test_systematic_definition():
    W = np.random.RandomState(1).normal(0, 1, 1000)
    a = resampling.ancestors(W, None, kind="systematic", u0=0.37)
    assert np.all(np.diff(a) >= 0)  # sorted ancestors
    counts = np.bincount(a, minlength=1000)
    w = weights.normalized_weights(W)
    assert np.all(np.abs(counts - 1000 * w) < 1.0 + 1e-9)  # |N_i - N w_i| < 1: the defining property


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = tuple(int(v) for v in rng.philox4x32_10(*ctr, *key))
        assert got == want


def test_normals_moments():
    z = rng.normals(3, np.arange(20000), 5, 4, rng.KIND_O_STEP)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    z2 = rng.normals(3, np.arange(20000), 5, 4, rng.KIND_MAXWELL_BOLTZMANN)
    assert abs(np.corrcoef(z.ravel(), z2.ravel())[0, 1]) < 0.02  # streams are independent


@pytest.mark.parametrize("n", [1, 5, 1023, 1024, 1025, 10000])
def test_canonical_statistics_equal_the_plain_definitions(n):
    """The fixed summation order of the resampling team (oracle.resampling.canonical_statistics) is just an order:
    nESS and the normaliser equal the textbook / reference forms (coddiwomple/utils.py:41-64, resamplers.py:107)."""
    from oracle import resampling, weights
    W = np.random.RandomState(n).normal(0.0, 2.0, n)
    wmin, sum_e, sum_e2, ness, mean = resampling.canonical_statistics(W)
    assert wmin == W.min()
    assert ness == pytest.approx(weights.nESS(W), rel=1e-12)
    assert mean == pytest.approx(-(np.logaddexp.reduce(-W) - np.log(n)), rel=1e-12, abs=1e-12)
    assert sum_e == pytest.approx(np.exp(wmin - W).sum(), rel=1e-13)
