"""fp64 numpy restatement of the reduced potential the reference obtains from OpenMM.  TEST INFRASTRUCTURE ONLY.

Reference behaviour restated: `OpenMMPDFState.reduced_potential` (coddiwomple/openmm/states.py:116-136) =
beta * U(x; lambda) of the OpenMM `System`; vacuum => no pV term (coddiwomple/openmm/coddiwomple.py:124-143).
The arithmetic itself lives in OpenMM 7.4 (un-vendored; version from the `openmmVersion="7.4"` header of the
System XML inside coddiwomple/data/perses_data/*.factory.pkl), so the potential is restated from the System XML:
every <Force> block is evaluated with the expression / standard OpenMM functional form it declares.

Everything is vectorised over particles: positions have shape (P, A, 3), fp64, nanometres; energies kJ/mol.
`energy_forces` also returns -dU/dx (analytic; checked against central differences in tests/test_oracle_potential.py).
"""
import gzip
import xml.etree.ElementTree as ET

import numpy as np

# constants implied by OpenMM 7.4 / simtk.unit
KB_KJ_PER_MOL_K = 1.3806488e-23 * 6.02214129e23 / 1000.0
ONE_4PI_EPS0 = 138.935456


def kT_kj_per_mol(temperature_kelvin=300.0):
    return KB_KJ_PER_MOL_K * temperature_kelvin


def load_system_xml(path):
    if str(path).endswith(".gz"):
        with gzip.open(path, "rb") as fh:
            return fh.read().decode("ascii")
    with open(path) as fh:
        return fh.read()


def _f(x):
    return float(x)


class XmlSystemOracle:
    """Evaluates U(x; lambda) and forces for an OpenMM System XML (vacuum, NoCutoff)."""

    def __init__(self, xml_text):
        root = ET.fromstring(xml_text)
        self.masses = np.array([_f(p.attrib["mass"]) for p in root.find("Particles")], dtype=np.float64)
        self.n_atoms = len(self.masses)
        cons = root.find("Constraints")
        self.constraints = [(int(c.attrib["p1"]), int(c.attrib["p2"]), _f(c.attrib["d"])) for c in (cons if cons is not None else [])]
        self.global_defaults = {}
        self.blocks = []
        for force in root.find("Forces"):
            gp = force.find("GlobalParameters")
            if gp is not None:
                for p in gp:
                    self.global_defaults.setdefault(p.attrib["name"], _f(p.attrib["default"]))
            self.blocks.append(self._parse_force(force))

    # ------------------------------------------------------------------ parsing
    def _parse_force(self, force):
        t = force.attrib["type"]
        energy = force.attrib.get("energy", "")
        if t == "HarmonicBondForce":
            b = force.find("Bonds")
            return ("hbond", dict(idx=np.array([[int(x.attrib["p1"]), int(x.attrib["p2"])] for x in b], dtype=int).reshape(-1, 2),
                                  d=np.array([_f(x.attrib["d"]) for x in b]), k=np.array([_f(x.attrib["k"]) for x in b])))
        if t == "HarmonicAngleForce":
            a = force.find("Angles")
            return ("hangle", dict(idx=np.array([[int(x.attrib["p1"]), int(x.attrib["p2"]), int(x.attrib["p3"])] for x in a], dtype=int).reshape(-1, 3),
                                   a=np.array([_f(x.attrib["a"]) for x in a]), k=np.array([_f(x.attrib["k"]) for x in a])))
        if t == "PeriodicTorsionForce":
            tt = force.find("Torsions")
            return ("ptorsion", dict(idx=np.array([[int(x.attrib[f"p{i}"]) for i in (1, 2, 3, 4)] for x in tt], dtype=int).reshape(-1, 4),
                                     n=np.array([_f(x.attrib["periodicity"]) for x in tt]),
                                     phase=np.array([_f(x.attrib["phase"]) for x in tt]), k=np.array([_f(x.attrib["k"]) for x in tt])))
        if t == "CustomBondForce":
            b = force.find("Bonds")
            names = [p.attrib["name"] for p in force.find("PerBondParameters")]
            idx = np.array([[int(x.attrib["p1"]), int(x.attrib["p2"])] for x in b], dtype=int).reshape(-1, 2)
            par = np.array([[_f(x.attrib[f"param{i + 1}"]) for i in range(len(names))] for x in b]).reshape(-1, len(names))
            if energy.startswith("(K/2)*(r-length)^2"):
                assert names == ["length1", "K1", "length2", "K2"], names
                return ("cbond", dict(idx=idx, par=par))
            if energy.startswith("U_electrostatics + U_sterics"):
                assert names[:5] == ["chargeProd", "sigmaA", "epsilonA", "sigmaB", "epsilonB"], names
                return ("cbond14", dict(idx=idx, par=par, names=names))
            raise NotImplementedError(f"CustomBondForce expression not recognised: {energy[:60]}")
        if t == "CustomAngleForce":
            a = force.find("Angles")
            names = [p.attrib["name"] for p in force.find("PerAngleParameters")]
            assert energy.startswith("(K/2)*(theta-theta0)^2") and names == ["theta0_1", "K_1", "theta0_2", "K_2"], (energy[:40], names)
            idx = np.array([[int(x.attrib["p1"]), int(x.attrib["p2"]), int(x.attrib["p3"])] for x in a], dtype=int).reshape(-1, 3)
            par = np.array([[_f(x.attrib[f"param{i + 1}"]) for i in range(4)] for x in a]).reshape(-1, 4)
            return ("cangle", dict(idx=idx, par=par))
        if t == "CustomTorsionForce":
            tt = force.find("Torsions")
            names = [p.attrib["name"] for p in force.find("PerTorsionParameters")]
            assert energy.startswith("(1-lambda_torsions)*U1 + lambda_torsions*U2") and names == ["periodicity1", "phase1", "K1", "periodicity2", "phase2", "K2"]
            idx = np.array([[int(x.attrib[f"p{i}"]) for i in (1, 2, 3, 4)] for x in tt], dtype=int).reshape(-1, 4)
            par = np.array([[_f(x.attrib[f"param{i + 1}"]) for i in range(6)] for x in tt]).reshape(-1, 6)
            return ("ctorsion", dict(idx=idx, par=par))
        if t == "CustomExternalForce":
            names = [p.attrib["name"] for p in force.find("PerParticleParameters")] if force.find("PerParticleParameters") is not None else []
            assert energy.replace(" ", "") == "(K/2)*((x-x0)^2+y^2+z^2)+U0".replace(" ", ""), energy
            gnames = [p.attrib["name"] for p in force.find("GlobalParameters")]
            pref = [g for g in gnames if g.endswith("_K")][0][:-2]
            parts = [int(p.attrib["index"]) for p in force.find("Particles")]
            return ("hoext", dict(atoms=np.array(parts, dtype=int), prefix=pref, names=names))
        if t == "NonbondedForce":
            assert force.attrib["method"] == "0", "only NoCutoff (vacuum) is restated"
            parts = force.find("Particles")
            q = np.array([_f(p.attrib["q"]) for p in parts]); sig = np.array([_f(p.attrib["sig"]) for p in parts]); eps = np.array([_f(p.attrib["eps"]) for p in parts])
            exc = force.find("Exceptions")
            eidx = np.array([[int(e.attrib["p1"]), int(e.attrib["p2"])] for e in exc], dtype=int).reshape(-1, 2)
            eq = np.array([_f(e.attrib["q"]) for e in exc]); esig = np.array([_f(e.attrib["sig"]) for e in exc]); eeps = np.array([_f(e.attrib["eps"]) for e in exc])
            poff = [(o.attrib["parameter"], int(o.attrib["particle"]), _f(o.attrib["q"]), _f(o.attrib["sig"]), _f(o.attrib["eps"]))
                    for o in (force.find("ParticleOffsets") if force.find("ParticleOffsets") is not None else [])]
            eoff = [(o.attrib["parameter"], int(o.attrib["exception"]), _f(o.attrib["q"]), _f(o.attrib["sig"]), _f(o.attrib["eps"]))
                    for o in (force.find("ExceptionOffsets") if force.find("ExceptionOffsets") is not None else [])]
            A = len(q)
            excepted = np.zeros((A, A), dtype=bool)
            for i, j in eidx:
                excepted[i, j] = excepted[j, i] = True
            iu, ju = np.triu_indices(A, 1)
            keep = ~excepted[iu, ju]
            return ("nonbonded", dict(q=q, sig=sig, eps=eps, eidx=eidx, eq=eq, esig=esig, eeps=eeps, poff=poff, eoff=eoff,
                                      pairs=np.stack([iu[keep], ju[keep]], 1)))
        if t == "CustomNonbondedForce":
            assert force.attrib["method"] == "0"
            assert energy.startswith("U_sterics;U_sterics = select(step(r - r_LJ)"), energy[:60]
            names = [p.attrib["name"] for p in force.find("PerParticleParameters")]
            assert names == ["sigmaA", "epsilonA", "sigmaB", "epsilonB", "unique_old", "unique_new"], names
            par = np.array([[_f(p.attrib[f"param{i + 1}"]) for i in range(6)] for p in force.find("Particles")])
            A = par.shape[0]
            excl = np.zeros((A, A), dtype=bool)
            for e in force.find("Exclusions"):
                i, j = int(e.attrib["p1"]), int(e.attrib["p2"])
                excl[i, j] = excl[j, i] = True
            groups = force.find("InteractionGroups")
            inter = np.zeros((A, A), dtype=bool)
            if groups is None or len(groups) == 0:
                inter[:] = True
            else:
                for g in groups:
                    s1 = [int(p.attrib["index"]) for p in g.find("Set1")]
                    s2 = [int(p.attrib["index"]) for p in g.find("Set2")]
                    for i in s1:
                        for j in s2:
                            if i != j:
                                inter[i, j] = inter[j, i] = True
            iu, ju = np.triu_indices(A, 1)
            keep = inter[iu, ju] & ~excl[iu, ju]
            return ("csterics", dict(par=par, pairs=np.stack([iu[keep], ju[keep]], 1)))
        if t in ("CMMotionRemover",):
            return ("noop", {})
        raise NotImplementedError(f"force type {t} is outside the vacuum hot path")

    # ------------------------------------------------------------------ helpers
    def lambdas(self, overrides=None):
        lam = dict(self.global_defaults)
        if overrides:
            for k, v in overrides.items():
                if k not in lam:
                    raise KeyError(f"unknown global parameter {k}")
                lam[k] = float(v)
        return lam

    @staticmethod
    def _bond(pos, idx, K, L, F):
        d = pos[:, idx[:, 0]] - pos[:, idx[:, 1]]
        r = np.sqrt((d * d).sum(-1))
        e = 0.5 * K * (r - L) ** 2
        if F is not None:
            g = (K * (r - L) / r)[..., None] * d
            np.subtract.at(F, (slice(None), idx[:, 0]), g)
            np.add.at(F, (slice(None), idx[:, 1]), g)
        return e.sum(-1)

    @staticmethod
    def _angle(pos, idx, K, T0, F):
        a = pos[:, idx[:, 0]] - pos[:, idx[:, 1]]
        b = pos[:, idx[:, 2]] - pos[:, idx[:, 1]]
        ra = np.sqrt((a * a).sum(-1)); rb = np.sqrt((b * b).sum(-1))
        c = np.clip((a * b).sum(-1) / (ra * rb), -1.0, 1.0)
        th = np.arccos(c)
        e = 0.5 * K * (th - T0) ** 2
        if F is not None:
            dedth = K * (th - T0)
            s = np.sqrt(np.maximum(1.0 - c * c, 1e-24))
            ah = a / ra[..., None]; bh = b / rb[..., None]
            fi = (dedth / (s * ra))[..., None] * (bh - c[..., None] * ah)
            fk = (dedth / (s * rb))[..., None] * (ah - c[..., None] * bh)
            np.add.at(F, (slice(None), idx[:, 0]), fi)
            np.add.at(F, (slice(None), idx[:, 2]), fk)
            np.subtract.at(F, (slice(None), idx[:, 1]), fi + fk)
        return e.sum(-1)

    @staticmethod
    def dihedral(pos, idx):
        """OpenMM sign convention (v0 = p1-p2, v1 = p3-p2, v2 = p3-p4; sign of v0.(v1 x v2))."""
        b1 = pos[:, idx[:, 1]] - pos[:, idx[:, 0]]
        b2 = pos[:, idx[:, 2]] - pos[:, idx[:, 1]]
        b3 = pos[:, idx[:, 3]] - pos[:, idx[:, 2]]
        n1 = np.cross(b1, b2); n2 = np.cross(b2, b3)
        rb2 = np.sqrt((b2 * b2).sum(-1))
        y = (np.cross(n1, n2) * b2).sum(-1) / rb2
        x = (n1 * n2).sum(-1)
        return np.arctan2(y, x), (b1, b2, b3, n1, n2, rb2)

    @classmethod
    def _torsion(cls, pos, idx, k, n, phase, F):
        phi, (b1, b2, b3, n1, n2, rb2) = cls.dihedral(pos, idx)
        e = k * (1.0 + np.cos(n * phi - phase))
        if F is not None:
            dedphi = -k * n * np.sin(n * phi - phase)
            n1sq = (n1 * n1).sum(-1); n2sq = (n2 * n2).sum(-1)
            fi = (dedphi * rb2 / n1sq)[..., None] * n1
            fl = (-dedphi * rb2 / n2sq)[..., None] * n2
            s = ((b1 * b2).sum(-1) / (rb2 * rb2))[..., None]
            t = ((b3 * b2).sum(-1) / (rb2 * rb2))[..., None]
            fj = -fi - s * fi + t * fl
            fk = -fl + s * fi - t * fl
            for col, f in zip(range(4), (fi, fj, fk, fl)):
                np.add.at(F, (slice(None), idx[:, col]), f)
        return e.sum(-1)

    @staticmethod
    def _coulomb_lj(pos, pairs, qq, sig, eps, F):
        d = pos[:, pairs[:, 0]] - pos[:, pairs[:, 1]]
        r2 = (d * d).sum(-1)
        r = np.sqrt(r2)
        sr6 = (sig * sig / r2) ** 3
        e = ONE_4PI_EPS0 * qq / r + 4.0 * eps * sr6 * (sr6 - 1.0)
        if F is not None:
            dedr = -ONE_4PI_EPS0 * qq / r2 + 4.0 * eps * (-12.0 * sr6 * sr6 + 6.0 * sr6) / r
            g = (dedr / r)[..., None] * d
            np.subtract.at(F, (slice(None), pairs[:, 0]), g)
            np.add.at(F, (slice(None), pairs[:, 1]), g)
        return e.sum(-1)

    @staticmethod
    def softcore_lj(r, sigma, epsilon, r_lj):
        """U_sterics of the perses soft-core LJ v2 expression and dU/dr; broadcasting over r (…, n)."""
        x = (sigma / r) ** 6
        e_plain = 4.0 * epsilon * x * (x - 1.0)
        d_plain = 4.0 * epsilon * (-12.0 * x * x + 6.0 * x) / r
        safe = np.where(r_lj > 0, r_lj, 1.0)
        xc = (sigma / safe) ** 6
        force = -4.0 * epsilon * (-12.0 * sigma ** 12 / safe ** 13 + 6.0 * sigma ** 6 / safe ** 7)
        ucut = 4.0 * epsilon * xc * (xc - 1.0)
        dr = r - safe
        e_quad = force * (dr * dr / 2.0 - dr) + ucut
        d_quad = force * (dr - 1.0)
        plain = (r >= r_lj) | (r_lj <= 0)
        return np.where(plain, e_plain, e_quad), np.where(plain, d_plain, d_quad)

    # ------------------------------------------------------------------ evaluation
    def energy_forces(self, positions, lambdas=None, want_forces=True, per_block=False):
        pos = np.asarray(positions, dtype=np.float64)
        single = pos.ndim == 2
        if single:
            pos = pos[None]
        lam = self.lambdas(lambdas)
        P = pos.shape[0]
        F = np.zeros_like(pos) if want_forces else None
        U = np.zeros(P)
        parts = []
        for kind, b in self.blocks:
            e = np.zeros(P)
            if kind == "hbond" and len(b["k"]):
                e = self._bond(pos, b["idx"], b["k"], b["d"], F)
            elif kind == "cbond" and len(b["par"]):
                l = lam["lambda_bonds"]; p = b["par"]
                K = (1 - l) * p[:, 1] + l * p[:, 3]; L = (1 - l) * p[:, 0] + l * p[:, 2]
                e = self._bond(pos, b["idx"], K, L, F)
            elif kind == "hangle" and len(b["k"]):
                e = self._angle(pos, b["idx"], b["k"], b["a"], F)
            elif kind == "cangle" and len(b["par"]):
                l = lam["lambda_angles"]; p = b["par"]
                K = (1.0 - l) * p[:, 1] + l * p[:, 3]; T0 = (1.0 - l) * p[:, 0] + l * p[:, 2]
                e = self._angle(pos, b["idx"], K, T0, F)
            elif kind == "ptorsion" and len(b["k"]):
                e = self._torsion(pos, b["idx"], b["k"], b["n"], b["phase"], F)
            elif kind == "ctorsion" and len(b["par"]):
                l = lam["lambda_torsions"]; p = b["par"]
                e = self._torsion(pos, b["idx"], (1 - l) * p[:, 2], p[:, 0], p[:, 1], F) + \
                    self._torsion(pos, b["idx"], l * p[:, 5], p[:, 3], p[:, 4], F)
            elif kind == "hoext":
                pre = b["prefix"]
                K, x0, U0 = lam[pre + "_K"], lam[pre + "_x0"], lam[pre + "_U0"]
                x = pos[:, b["atoms"]]
                d = x.copy(); d[..., 0] -= x0
                e = (0.5 * K * (d * d).sum(-1) + U0).sum(-1)
                if F is not None:
                    np.subtract.at(F, (slice(None), b["atoms"]), K * d)
            elif kind == "nonbonded":
                q = b["q"].copy(); sig = b["sig"].copy(); eps = b["eps"].copy()
                for (name, i, dq, ds, de) in b["poff"]:
                    q[i] += lam[name] * dq; sig[i] += lam[name] * ds; eps[i] += lam[name] * de
                eq = b["eq"].copy(); esig = b["esig"].copy(); eeps = b["eeps"].copy()
                for (name, i, dq, ds, de) in b["eoff"]:
                    eq[i] += lam[name] * dq; esig[i] += lam[name] * ds; eeps[i] += lam[name] * de
                pr = b["pairs"]
                if len(pr):
                    e = e + self._coulomb_lj(pos, pr, q[pr[:, 0]] * q[pr[:, 1]], 0.5 * (sig[pr[:, 0]] + sig[pr[:, 1]]),
                                             np.sqrt(eps[pr[:, 0]] * eps[pr[:, 1]]), F)
                if len(b["eidx"]):
                    e = e + self._coulomb_lj(pos, b["eidx"], eq, esig, eeps, F)
            elif kind == "csterics" and len(b["pairs"]):
                p = b["par"]; pr = b["pairs"]; i, j = pr[:, 0], pr[:, 1]
                uo = np.maximum(p[i, 4], p[j, 4]); un = np.maximum(p[i, 5], p[j, 5])
                core = ((p[i, 4] + p[j, 4] + p[i, 5] + p[j, 5]) == 0).astype(float)
                ls = core * lam["lambda_sterics_core"] + un * lam["lambda_sterics_insert"] + uo * lam["lambda_sterics_delete"]
                epsA = np.sqrt(p[i, 1] * p[j, 1]); epsB = np.sqrt(p[i, 3] * p[j, 3])
                sigA = 0.5 * (p[i, 0] + p[j, 0]); sigB = 0.5 * (p[i, 2] + p[j, 2])
                eps = (1 - ls) * epsA + ls * epsB; sig = (1 - ls) * sigA + ls * sigB
                ldep = un * (1.0 - lam["lambda_sterics_insert"]) + uo * lam["lambda_sterics_delete"]
                rlj = lam["softcore_alpha"] * ((26.0 / 7.0) * sig ** 6 * ldep) ** (1.0 / 6.0)
                d = pos[:, i] - pos[:, j]
                r = np.sqrt((d * d).sum(-1))
                ee, dedr = self.softcore_lj(r, sig, eps, rlj)
                e = ee.sum(-1)
                if F is not None:
                    g = (dedr / r)[..., None] * d
                    np.subtract.at(F, (slice(None), i), g)
                    np.add.at(F, (slice(None), j), g)
            elif kind == "cbond14" and len(b["par"]):
                p = b["par"]; names = b["names"]; i, j = b["idx"][:, 0], b["idx"][:, 1]
                uo = p[:, names.index("unique_old")]; un = p[:, names.index("unique_new")]
                new = (un == 1).astype(float); old = (uo == 1).astype(float)
                ls = new * lam["lambda_sterics_insert"] + old * lam["lambda_sterics_delete"]
                eps = (1 - ls) * p[:, 2] + ls * p[:, 4]; sig = (1 - ls) * p[:, 1] + ls * p[:, 3]
                ldep = new * (1.0 - lam["lambda_sterics_insert"]) + old * lam["lambda_sterics_delete"]
                rlj = lam["softcore_alpha"] * ((26.0 / 7.0) * sig ** 6 * ldep) ** (1.0 / 6.0)
                d = pos[:, i] - pos[:, j]
                r = np.sqrt((d * d).sum(-1))
                ee, dedr = self.softcore_lj(r, sig, eps, rlj)
                scale = lam["lambda_electrostatics_insert"] * un + uo * (1 - lam["lambda_electrostatics_delete"])
                ee = ee + scale * ONE_4PI_EPS0 * p[:, 0] / r
                dedr = dedr - scale * ONE_4PI_EPS0 * p[:, 0] / (r * r)
                e = ee.sum(-1)
                if F is not None:
                    g = (dedr / r)[..., None] * d
                    np.subtract.at(F, (slice(None), i), g)
                    np.add.at(F, (slice(None), j), g)
            parts.append((kind, e))
            U = U + e
        if single:
            U = U[0]
            F = F[0] if F is not None else None
        if per_block:
            return U, F, parts
        return U, F

    def potential_energy(self, positions, lambdas=None):
        return self.energy_forces(positions, lambdas, want_forces=False)[0]

    def reduced_potential(self, positions, lambdas=None, temperature_kelvin=300.0):
        """beta * U — restates coddiwomple/openmm/states.py:116-136 for pressure=None."""
        return self.potential_energy(positions, lambdas) / kT_kj_per_mol(temperature_kelvin)
