"""Test systems as OpenMM-7.4-style System XML (the format the reference's fixtures use), generated without OpenMM.

* `harmonic_oscillator_xml` — the system of `openmmtools.testsystems.HarmonicOscillator` the reference tests use
  (coddiwomple/tests/utils.py:27-72; parameter names confirmed by troubleshooting/test_suite.ipynb JSON 41-45):
  one particle, CustomExternalForce (K/2)((x-x0)^2+y^2+z^2)+U0 with globals testsystems_HarmonicOscillator_{K,x0,U0}.
* `alanine_dipeptide_shaped_xml` — BASELINE.json config 2.  openmmtools' AlanineDipeptideVacuum (amber96) cannot be
  obtained offline, so this is a SYNTHETIC 22-atom system with the real ACE-ALA-NME topology (21 bonds of which the
  12 X-H are constraints, 36 angles, 41 proper torsions, 231 atom pairs) and seeded amber/GAFF-like parameters, cast
  in the perses hybrid form of the reference's fixtures: the ALA backbone is an alchemical "core" whose charges /
  LJ / bonded terms interpolate between two end states, the three CB hydrogens are "unique old" atoms that are
  decoupled through the soft-core LJ v2 expression, ACE and NME are the unperturbed environment.  Parity for it is
  unpinned by the reference (no artefact exists); it is checked against the fp64 oracle only.
"""
import numpy as np

from .system import kT_in_md_units

HO_PREFIX = "testsystems_HarmonicOscillator"

RELATIVE_ALCHEMICAL_PARAMETERS = (
    "lambda_sterics_core", "lambda_electrostatics_core", "lambda_sterics_insert", "lambda_sterics_delete",
    "lambda_electrostatics_insert", "lambda_electrostatics_delete", "lambda_bonds", "lambda_angles", "lambda_torsions")


def harmonic_oscillator_parameters(temperature=300.0, sigma_nm=0.1, mass=39.948):
    """K, period, collision_rate, timestep exactly as coddiwomple/tests/utils.py:55-60 derives them (MD units)."""
    kT = kT_in_md_units(temperature)
    K = kT / sigma_nm ** 2
    period = np.sqrt(mass / K)
    return dict(K=K, kT=kT, mass=mass, period=period, timestep=period / 20.0, collision_rate=1.0 / period, sigma=sigma_nm)


def harmonic_oscillator_xml(K=None, mass=39.948, temperature=300.0, sigma_nm=0.1):
    if K is None:
        K = harmonic_oscillator_parameters(temperature, sigma_nm, mass)["K"]
    return f"""<System openmmVersion="7.4" type="System" version="1">
	<PeriodicBoxVectors><A x="2" y="0" z="0"/><B x="0" y="2" z="0"/><C x="0" y="0" z="2"/></PeriodicBoxVectors>
	<Particles><Particle mass="{mass!r}"/></Particles>
	<Constraints/>
	<Forces>
		<Force energy="(K/2)*((x-x0)^2+y^2+z^2)+U0" forceGroup="0" type="CustomExternalForce" version="1">
			<PerParticleParameters/>
			<GlobalParameters>
				<Parameter default="{K!r}" name="{HO_PREFIX}_K"/>
				<Parameter default="0" name="{HO_PREFIX}_x0"/>
				<Parameter default="0" name="{HO_PREFIX}_U0"/>
			</GlobalParameters>
			<Particles><Particle index="0"/></Particles>
		</Force>
	</Forces>
</System>"""


def harmonic_parameter_sequence(n_states, temperature=300.0, sigma_nm=0.1):
    """x0: 0 -> 2 sigma, U0: 0 -> 1 kT, linear (coddiwomple/tests/utils.py:63-67): Delta F = 1 kT exactly."""
    kT = kT_in_md_units(temperature)
    return [{f"{HO_PREFIX}_x0": 2.0 * sigma_nm * q, f"{HO_PREFIX}_U0": kT * q} for q in np.linspace(0.0, 1.0, n_states)]


def relative_alchemical_sequence(n_states):
    """The lambda protocol of the reference's SMC test (coddiwomple/tests/legacy_tests.py:58-78)."""
    seq = []
    for q in np.linspace(0.0, 1.0, n_states):
        q = float(q)
        seq.append({
            "lambda_sterics_core": q, "lambda_electrostatics_core": q,
            "lambda_sterics_insert": 2.0 * q if q < 0.5 else 1.0,
            "lambda_sterics_delete": 0.0 if q < 0.5 else 2.0 * (q - 0.5),
            "lambda_electrostatics_insert": 0.0 if q < 0.5 else 2.0 * (q - 0.5),
            "lambda_electrostatics_delete": 2.0 * q if q < 0.5 else 1.0,
            "lambda_bonds": q, "lambda_angles": q, "lambda_torsions": q})
    return seq


# ---------------------------------------------------------------------------------------------- ACE-ALA-NME shape
_ALA_NAMES = ["HH31", "CH3", "HH32", "HH33", "C", "O", "N", "H", "CA", "HA", "CB", "HB1", "HB2", "HB3", "C", "O", "N", "H", "CH3", "HH31", "HH32", "HH33"]
_ALA_TYPES = ["HC", "CT", "HC", "HC", "C", "O", "N", "HN", "CT", "H1", "CT", "HC", "HC", "HC", "C", "O", "N", "HN", "CT", "H1", "H1", "H1"]
_ALA_BONDS = [(1, 0), (1, 2), (1, 3), (1, 4), (4, 5), (4, 6), (6, 7), (6, 8), (8, 9), (8, 10), (10, 11), (10, 12), (10, 13), (8, 14), (14, 15),
              (14, 16), (16, 17), (16, 18), (18, 19), (18, 20), (18, 21)]
# z-matrix: atom -> (bonded to, angle with, dihedral with, r nm, theta deg, phi deg)
_ALA_ZMAT = {
    1: None, 4: (1, None, None, 0.1522, None, None), 0: (1, 4, None, 0.1090, 109.5, None),
    2: (1, 4, 0, 0.1090, 109.5, 120.0), 3: (1, 4, 0, 0.1090, 109.5, -120.0),
    5: (4, 1, 0, 0.1229, 120.4, 10.0), 6: (4, 1, 5, 0.1335, 116.6, 180.0),
    7: (6, 4, 5, 0.1010, 119.8, 180.0), 8: (6, 4, 5, 0.1449, 121.9, 0.0),
    9: (8, 6, 4, 0.1090, 109.5, 45.0), 10: (8, 6, 9, 0.1526, 109.7, 120.0), 14: (8, 6, 9, 0.1522, 110.1, -120.0),
    11: (10, 8, 6, 0.1090, 109.5, 60.0), 12: (10, 8, 11, 0.1090, 109.5, 120.0), 13: (10, 8, 11, 0.1090, 109.5, -120.0),
    15: (14, 8, 6, 0.1229, 120.4, -30.0), 16: (14, 8, 15, 0.1335, 116.6, 180.0),
    17: (16, 14, 15, 0.1010, 119.8, 180.0), 18: (16, 14, 15, 0.1449, 121.9, 0.0),
    19: (18, 16, 14, 0.1090, 109.5, 60.0), 20: (18, 16, 19, 0.1090, 109.5, 120.0), 21: (18, 16, 19, 0.1090, 109.5, -120.0),
}
_MASS = {"H": 1.008, "C": 12.01, "N": 14.01, "O": 16.0}
_CHARGE = {"HC": 0.06, "CT": -0.12, "C": 0.60, "O": -0.57, "N": -0.42, "HN": 0.27, "H1": 0.09}
_LJ = {"HC": (0.2650, 0.0657), "H1": (0.2471, 0.0657), "HN": (0.1069, 0.0657), "CT": (0.3400, 0.4577), "C": (0.3400, 0.3598),
       "N": (0.3250, 0.7113), "O": (0.2960, 0.8786)}


def _place(a, b, c, r, theta, phi):
    """Position at distance r from a, angle theta (a is the vertex, towards b), dihedral phi about a-b relative to c."""
    bc = a - b
    bc /= np.linalg.norm(bc)
    n = np.cross(b - c, bc)
    n /= np.linalg.norm(n)
    m = np.cross(n, bc)
    d = np.array([-r * np.cos(theta), r * np.sin(theta) * np.cos(phi), r * np.sin(theta) * np.sin(phi)])
    return a + d[0] * bc + d[1] * m + d[2] * n


def alanine_dipeptide_shaped_geometry():
    x = np.zeros((22, 3))
    order = [1, 4, 0, 2, 3, 5, 6, 7, 8, 9, 10, 14, 11, 12, 13, 15, 16, 17, 18, 19, 20, 21]
    for at in order:
        z = _ALA_ZMAT[at]
        if z is None:
            continue
        b, a, d, r, th, ph = z
        if a is None:
            x[at] = x[b] + np.array([r, 0.0, 0.0])
        elif d is None:
            th = np.deg2rad(th)
            u = x[a] - x[b]
            u /= np.linalg.norm(u)
            perp = np.array([0.0, 1.0, 0.0])
            x[at] = x[b] + r * (np.cos(th) * u + np.sin(th) * perp)
        else:
            x[at] = _place(x[b], x[a], x[d], r, np.deg2rad(th), np.deg2rad(ph))
    return x


def alanine_dipeptide_shaped_xml(seed=2):
    """Returns (xml, positions_nm[22,3]).  See the module docstring for what is real and what is synthetic."""
    rng = np.random.RandomState(seed)
    A = 22
    x = alanine_dipeptide_shaped_geometry()
    elem = [t[0] for t in _ALA_TYPES]
    masses = [_MASS[e] for e in elem]
    nbrs = [set() for _ in range(A)]
    for i, j in _ALA_BONDS:
        nbrs[i].add(j)
        nbrs[j].add(i)
    core = {6, 7, 8, 9, 10, 14, 15}
    uold = {11, 12, 13}
    env = set(range(A)) - core - uold

    def dist(i, j):
        return float(np.linalg.norm(x[i] - x[j]))

    def angle(i, j, k):
        a, b = x[i] - x[j], x[k] - x[j]
        return float(np.arccos(np.dot(a, b) / np.linalg.norm(a) / np.linalg.norm(b)))

    constraints = [(i, j, dist(i, j)) for i, j in _ALA_BONDS if "H" in (elem[i], elem[j])]
    heavy_bonds = [(i, j) for i, j in _ALA_BONDS if "H" not in (elem[i], elem[j])]
    angles = sorted({(min(i, k), j, max(i, k)) for j in range(A) for i in nbrs[j] for k in nbrs[j] if i != k})
    torsions = []
    for j, k in _ALA_BONDS:
        for i in sorted(nbrs[j] - {k}):
            for l in sorted(nbrs[k] - {j}):
                torsions.append((i, j, k, l))
    one2 = {(min(i, j), max(i, j)) for i, j in _ALA_BONDS}
    one3 = {(a[0], a[2]) for a in angles}
    one4 = {(min(t[0], t[3]), max(t[0], t[3])) for t in torsions} - one2 - one3

    # ---- charges (two end states for the core), neutralised
    qA = np.array([_CHARGE[t] for t in _ALA_TYPES]) + rng.normal(0, 0.01, A)
    qA -= qA.mean()
    qB = qA.copy()
    for a in core:
        qB[a] = 0.8 * qA[a] + rng.normal(0, 0.02)
    qB[list(core)] -= (qB.sum() - qA.sum() + qA[list(uold)].sum()) / len(core)  # end state B is neutral without the deleted atoms' charge
    sig = np.array([_LJ[t][0] for t in _ALA_TYPES])
    eps = np.array([_LJ[t][1] for t in _ALA_TYPES])
    sigB, epsB = sig.copy(), eps.copy()
    for a in core:
        sigB[a] = sig[a] * (1.0 + rng.uniform(-0.05, 0.05))
        epsB[a] = eps[a] * (1.0 + rng.uniform(-0.2, 0.2))
    for a in uold:
        epsB[a] = 0.0

    out = ['<System openmmVersion="7.4" type="System" version="1">',
           '\t<PeriodicBoxVectors><A x="2" y="0" z="0"/><B x="0" y="2" z="0"/><C x="0" y="0" z="2"/></PeriodicBoxVectors>',
           "\t<Particles>"] + [f'\t\t<Particle mass="{m!r}"/>' for m in masses] + ["\t</Particles>", "\t<Constraints>"] + \
          [f'\t\t<Constraint d="{d!r}" p1="{i}" p2="{j}"/>' for i, j, d in constraints] + ["\t</Constraints>", "\t<Forces>"]

    def kbond(i, j):
        return {"CC": 265265.6, "CN": 410032.0, "CO": 476976.0}.get("".join(sorted(elem[i] + elem[j])), 300000.0) * (1 + rng.uniform(-0.05, 0.05))

    # bonds: core-core heavy bonds interpolate (CustomBondForce), the rest are plain harmonic
    cb, hb = [], []
    for i, j in heavy_bonds:
        k, l = kbond(i, j), dist(i, j)
        if i in core and j in core:
            cb.append((i, j, l, k, l * (1 + rng.uniform(-0.01, 0.01)), k * (1 + rng.uniform(-0.1, 0.1))))
        else:
            hb.append((i, j, l, k))
    out += ['\t\t<Force energy="(K/2)*(r-length)^2;K = (1-lambda_bonds)*K1 + lambda_bonds*K2;length = (1-lambda_bonds)*length1 + lambda_bonds*length2;" forceGroup="0" type="CustomBondForce" usesPeriodic="0" version="3">',
            '\t\t\t<PerBondParameters><Parameter name="length1"/><Parameter name="K1"/><Parameter name="length2"/><Parameter name="K2"/></PerBondParameters>',
            '\t\t\t<GlobalParameters><Parameter default="0" name="lambda_bonds"/></GlobalParameters>', "\t\t\t<Bonds>"] + \
           [f'\t\t\t\t<Bond p1="{i}" p2="{j}" param1="{l1!r}" param2="{k1!r}" param3="{l2!r}" param4="{k2!r}"/>' for i, j, l1, k1, l2, k2 in cb] + \
           ["\t\t\t</Bonds>", "\t\t</Force>", '\t\t<Force forceGroup="0" type="HarmonicBondForce" usesPeriodic="0" version="2">', "\t\t\t<Bonds>"] + \
           [f'\t\t\t\t<Bond d="{l!r}" k="{k!r}" p1="{i}" p2="{j}"/>' for i, j, l, k in hb] + ["\t\t\t</Bonds>", "\t\t</Force>"]

    ca, ha = [], []
    for i, j, k in angles:
        kk = (292.88 if "H" in (elem[i], elem[k]) else 585.76) * (1 + rng.uniform(-0.1, 0.1))
        th = angle(i, j, k)
        if {i, j, k} <= core:
            ca.append((i, j, k, th, kk, th + rng.uniform(-0.02, 0.02), kk * (1 + rng.uniform(-0.1, 0.1))))
        else:
            ha.append((i, j, k, th, kk))
    out += ['\t\t<Force energy="(K/2)*(theta-theta0)^2;K = (1.0-lambda_angles)*K_1 + lambda_angles*K_2;theta0 = (1.0-lambda_angles)*theta0_1 + lambda_angles*theta0_2;" forceGroup="0" type="CustomAngleForce" usesPeriodic="0" version="3">',
            '\t\t\t<PerAngleParameters><Parameter name="theta0_1"/><Parameter name="K_1"/><Parameter name="theta0_2"/><Parameter name="K_2"/></PerAngleParameters>',
            '\t\t\t<GlobalParameters><Parameter default="0" name="lambda_angles"/></GlobalParameters>', "\t\t\t<Angles>"] + \
           [f'\t\t\t\t<Angle p1="{i}" p2="{j}" p3="{k}" param1="{t1!r}" param2="{k1!r}" param3="{t2!r}" param4="{k2!r}"/>' for i, j, k, t1, k1, t2, k2 in ca] + \
           ["\t\t\t</Angles>", "\t\t</Force>", '\t\t<Force forceGroup="0" type="HarmonicAngleForce" usesPeriodic="0" version="2">', "\t\t\t<Angles>"] + \
           [f'\t\t\t\t<Angle a="{t!r}" k="{k_!r}" p1="{i}" p2="{j}" p3="{k}"/>' for i, j, k, t, k_ in ha] + ["\t\t\t</Angles>", "\t\t</Force>"]

    ct, pt = [], []
    for (i, j, k, l) in torsions:
        amide = {elem[j], elem[k]} == {"C", "N"}
        n, ph, kk = (2, np.pi, 10.46 * (1 + rng.uniform(-0.1, 0.1))) if amide else (3, 0.0, 0.65 * (1 + rng.uniform(-0.3, 0.3)))
        if {i, j, k, l} <= core:
            ct.append((i, j, k, l, n, ph, kk, int(rng.choice([1, 2, 3])), float(rng.choice([0.0, np.pi])), kk * rng.uniform(0.5, 1.5)))
        else:
            pt.append((i, j, k, l, n, ph, kk))
    out += ['\t\t<Force energy="(1-lambda_torsions)*U1 + lambda_torsions*U2;U1 = K1*(1+cos(periodicity1*theta-phase1));U2 = K2*(1+cos(periodicity2*theta-phase2));" forceGroup="0" type="CustomTorsionForce" usesPeriodic="0" version="3">',
            '\t\t\t<PerTorsionParameters><Parameter name="periodicity1"/><Parameter name="phase1"/><Parameter name="K1"/><Parameter name="periodicity2"/><Parameter name="phase2"/><Parameter name="K2"/></PerTorsionParameters>',
            '\t\t\t<GlobalParameters><Parameter default="0" name="lambda_torsions"/></GlobalParameters>', "\t\t\t<Torsions>"] + \
           [f'\t\t\t\t<Torsion p1="{i}" p2="{j}" p3="{k}" p4="{l}" param1="{n1}" param2="{p1!r}" param3="{k1!r}" param4="{n2}" param5="{p2!r}" param6="{k2!r}"/>'
            for i, j, k, l, n1, p1, k1, n2, p2, k2 in ct] + \
           ["\t\t\t</Torsions>", "\t\t</Force>", '\t\t<Force forceGroup="0" type="PeriodicTorsionForce" usesPeriodic="0" version="2">', "\t\t\t<Torsions>"] + \
           [f'\t\t\t\t<Torsion k="{kk!r}" p1="{i}" p2="{j}" p3="{k}" p4="{l}" periodicity="{n}" phase="{ph!r}"/>' for i, j, k, l, n, ph, kk in pt] + \
           ["\t\t\t</Torsions>", "\t\t</Force>"]

    # ---- NonbondedForce: charges with offsets; LJ only between environment atoms (alchemical sterics live in the custom force)
    nb_eps = np.where([a in env for a in range(A)], eps, 0.0)
    out += ['\t\t<Force alpha="0" cutoff="1" dispersionCorrection="0" ewaldTolerance=".0005" forceGroup="0" method="0" rfDielectric="78.3" switchingDistance="-1" type="NonbondedForce" useSwitchingFunction="0" version="3">',
            "\t\t\t<GlobalParameters>"] + [f'\t\t\t\t<Parameter default="0" name="{n}"/>' for n in
                                            ("lambda_electrostatics_core", "lambda_sterics_core", "lambda_electrostatics_delete", "lambda_electrostatics_insert")] + \
           ["\t\t\t</GlobalParameters>", "\t\t\t<ParticleOffsets>"]
    for a in range(A):
        if a in core:
            out.append(f'\t\t\t\t<Offset eps="0" parameter="lambda_electrostatics_core" particle="{a}" q="{float(qB[a] - qA[a])!r}" sig="0"/>')
        elif a in uold:
            out.append(f'\t\t\t\t<Offset eps="0" parameter="lambda_electrostatics_delete" particle="{a}" q="{float(-qA[a])!r}" sig="0"/>')
    out += ["\t\t\t</ParticleOffsets>", "\t\t\t<ExceptionOffsets>"]
    exceptions, exc_offsets = [], []
    for (i, j) in sorted(one2 | one3):
        exceptions.append((i, j, 0.0, 1.0, 0.0))
    for (i, j) in sorted(one4):
        qq, s, e = qA[i] * qA[j] / 1.2, 0.5 * (sig[i] + sig[j]), np.sqrt(eps[i] * eps[j]) / 2.0
        k = len(exceptions)
        exceptions.append((i, j, float(qq), float(s), float(e)))
        if (i in core or j in core) and not (i in uold or j in uold):
            qqB, sB, eB = qB[i] * qB[j] / 1.2, 0.5 * (sigB[i] + sigB[j]), np.sqrt(epsB[i] * epsB[j]) / 2.0
            exc_offsets.append(("lambda_electrostatics_core", k, float(qqB - qq), 0.0, 0.0))
            exc_offsets.append(("lambda_sterics_core", k, 0.0, float(sB - s), float(eB - e)))
    out += [f'\t\t\t\t<Offset eps="{e!r}" exception="{k}" parameter="{n}" q="{q!r}" sig="{s!r}"/>' for n, k, q, s, e in exc_offsets]
    out += ["\t\t\t</ExceptionOffsets>", "\t\t\t<Particles>"] + \
           [f'\t\t\t\t<Particle eps="{float(nb_eps[a])!r}" q="{float(qA[a])!r}" sig="{float(sig[a])!r}"/>' for a in range(A)] + \
           ["\t\t\t</Particles>", "\t\t\t<Exceptions>"] + \
           [f'\t\t\t\t<Exception eps="{e!r}" p1="{i}" p2="{j}" q="{q!r}" sig="{s!r}"/>' for i, j, q, s, e in exceptions] + ["\t\t\t</Exceptions>", "\t\t</Force>"]

    # ---- CustomNonbondedForce: perses soft-core LJ v2 (expression copied from the form in the reference's fixture XML)
    expr = ("U_sterics;U_sterics = select(step(r - r_LJ), 4*epsilon*x*(x-1.0), U_sterics_quad);U_sterics_quad = Force*(((r - r_LJ)^2)/2 - (r - r_LJ)) + U_sterics_cut;"
            "U_sterics_cut = 4*epsilon*((sigma/r_LJ)^6)*(((sigma/r_LJ)^6) - 1.0);Force = -4*epsilon*((-12*sigma^12)/(r_LJ^13) + (6*sigma^6)/(r_LJ^7));x = (sigma/r)^6;"
            "r_LJ = softcore_alpha*((26/7)*(sigma^6)*lambda_sterics_deprecated)^(1/6);lambda_sterics_deprecated = new_interaction*(1.0 - lambda_sterics_insert) + old_interaction*lambda_sterics_delete;"
            "epsilon = (1-lambda_sterics)*epsilonA + lambda_sterics*epsilonB;reff_sterics = sigma*((softcore_alpha*lambda_alpha + (r/sigma)^6))^(1/6);sigma = (1-lambda_sterics)*sigmaA + lambda_sterics*sigmaB;"
            "lambda_alpha = new_interaction*(1-lambda_sterics_insert) + old_interaction*lambda_sterics_delete;lambda_sterics = core_interaction*lambda_sterics_core + new_interaction*lambda_sterics_insert + old_interaction*lambda_sterics_delete;"
            "core_interaction = delta(unique_old1+unique_old2+unique_new1+unique_new2);new_interaction = max(unique_new1, unique_new2);old_interaction = max(unique_old1, unique_old2);"
            "epsilonA = sqrt(epsilonA1*epsilonA2);epsilonB = sqrt(epsilonB1*epsilonB2);sigmaA = 0.5*(sigmaA1 + sigmaA2);sigmaB = 0.5*(sigmaB1 + sigmaB2);")
    out += [f'\t\t<Force cutoff="1" energy="{expr}" forceGroup="0" method="0" switchingDistance="-1" type="CustomNonbondedForce" useLongRangeCorrection="0" useSwitchingFunction="0" version="2">',
            '\t\t\t<PerParticleParameters><Parameter name="sigmaA"/><Parameter name="epsilonA"/><Parameter name="sigmaB"/><Parameter name="epsilonB"/><Parameter name="unique_old"/><Parameter name="unique_new"/></PerParticleParameters>',
            "\t\t\t<GlobalParameters>", '\t\t\t\t<Parameter default=".85" name="softcore_alpha"/>'] + \
           [f'\t\t\t\t<Parameter default="0" name="{n}"/>' for n in ("lambda_sterics_core", "lambda_electrostatics_core", "lambda_sterics_insert", "lambda_sterics_delete")] + \
           ["\t\t\t</GlobalParameters>", "\t\t\t<Particles>"] + \
           [f'\t\t\t\t<Particle param1="{float(sig[a])!r}" param2="{float(eps[a])!r}" param3="{float(sigB[a])!r}" param4="{float(epsB[a])!r}" param5="{1 if a in uold else 0}" param6="0"/>'
            for a in range(A)] + ["\t\t\t</Particles>", "\t\t\t<Exclusions>"] + \
           [f'\t\t\t\t<Exclusion p1="{i}" p2="{j}"/>' for i, j, *_ in exceptions] + ["\t\t\t</Exclusions>", "\t\t\t<Functions/>", "\t\t\t<InteractionGroups>"]
    for s1, s2 in ((uold, core), (uold, env), (uold, uold), (core, env), (core, core)):
        out += ["\t\t\t\t<InteractionGroup>", "\t\t\t\t\t<Set1>"] + [f'\t\t\t\t\t\t<Particle index="{a}"/>' for a in sorted(s1)] + \
               ["\t\t\t\t\t</Set1>", "\t\t\t\t\t<Set2>"] + [f'\t\t\t\t\t\t<Particle index="{a}"/>' for a in sorted(s2)] + ["\t\t\t\t\t</Set2>", "\t\t\t\t</InteractionGroup>"]
    out += ["\t\t\t</InteractionGroups>", "\t\t</Force>", "\t</Forces>", "</System>"]
    return "\n".join(out), x


def alanine_dipeptide_shaped_ensemble(n_particles, seed=2, jitter_nm=0.002):
    """Synthetic initial particles: the built geometry plus small Gaussian displacements (fp32-representable)."""
    _, x = alanine_dipeptide_shaped_xml(seed)
    rng = np.random.RandomState(seed + 1000)
    return (x[None] + rng.normal(0.0, jitter_nm, (n_particles,) + x.shape)).astype(np.float32)
