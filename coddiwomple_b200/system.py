"""Lower an OpenMM `System` (as serialised XML) to the term tables `cdw_system_create` takes.

The reference hands an `openmm.System` to `OpenMMPDFState` (coddiwomple/openmm/states.py:50-86), which lets OpenMM
evaluate it.  Here the same System — the plain OpenMM-7.4 XML that perses embeds in
coddiwomple/data/perses_data/*.factory.pkl, or `XmlSerializer.serialize(system)` when OpenMM is installed — is
lowered once on the host to struct-of-arrays tables in which every alchemical parameter is affine in one global
parameter (include/cdw.h: cdw_system_desc).  Only vacuum (NoCutoff) systems are in scope.
"""
import gzip
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

# OpenMM 7.4 / simtk.unit constants (md unit system)
KB_KJ_PER_MOL_K = 1.3806488e-23 * 6.02214129e23 / 1000.0
ONE_4PI_EPS0 = 138.935456


def kT_in_md_units(temperature_kelvin=300.0):
    """kB*N_A*T in kJ/mol."""
    return KB_KJ_PER_MOL_K * float(temperature_kelvin)


def _i32(rows, width):
    return np.ascontiguousarray(np.array(rows, dtype=np.int32).reshape(-1, width) if width > 1 else np.array(rows, dtype=np.int32).reshape(-1))


def _f64(rows, width):
    return np.ascontiguousarray(np.array(rows, dtype=np.float64).reshape(-1, width) if width > 1 else np.array(rows, dtype=np.float64).reshape(-1))


@dataclass
class SystemSpec:
    """Host-side term tables of one replica (see include/cdw.h for the meaning of every array)."""
    masses: np.ndarray
    lambda_names: list
    lambda_defaults: np.ndarray
    bond_atoms: np.ndarray
    bond_slot: np.ndarray
    bond_par: np.ndarray
    angle_atoms: np.ndarray
    angle_slot: np.ndarray
    angle_par: np.ndarray
    torsion_atoms: np.ndarray
    torsion_slot: np.ndarray
    torsion_n: np.ndarray
    torsion_par: np.ndarray
    nb_atom_par: np.ndarray
    atom_off_idx: np.ndarray
    atom_off_par: np.ndarray
    exc_par: np.ndarray
    exc_off_idx: np.ndarray
    exc_off_par: np.ndarray
    pair_atoms: np.ndarray
    pair_nb: np.ndarray
    pair_sc_class: np.ndarray
    pair_sc_par: np.ndarray
    sc_slots: tuple
    ext_atom: np.ndarray
    ext_slot: np.ndarray
    ext_par: np.ndarray
    cons_atoms: np.ndarray
    cons_dist: np.ndarray
    xml: str = field(default=None, repr=False)

    @property
    def n_atoms(self):
        return int(self.masses.shape[0])

    @property
    def n_lambda(self):
        return len(self.lambda_names)

    def lambda_vector(self, parameters=None):
        """Dense lambda vector for a `{name: value}` dict (missing names keep their XML defaults)."""
        lam = self.lambda_defaults.copy()
        if parameters:
            for k, v in parameters.items():
                if k not in self.lambda_names:
                    raise KeyError(f"unknown global parameter {k!r}; known: {self.lambda_names}")
                lam[self.lambda_names.index(k)] = float(v)
        return lam

    def counts(self):
        return dict(atoms=self.n_atoms, bonds=len(self.bond_slot), angles=len(self.angle_slot), torsions=len(self.torsion_slot),
                    pairs=len(self.pair_nb), constraints=len(self.cons_dist), ext=len(self.ext_atom), n_lambda=self.n_lambda)


def load_xml(path):
    if str(path).endswith(".gz"):
        with gzip.open(path, "rb") as fh:
            return fh.read().decode("ascii")
    with open(path) as fh:
        return fh.read()


def spec_from_xml(xml_text):
    root = ET.fromstring(xml_text)
    masses = np.array([float(p.attrib["mass"]) for p in root.find("Particles")], dtype=np.float64)
    A = len(masses)
    cons = root.find("Constraints")
    cons_rows = [(int(c.attrib["p1"]), int(c.attrib["p2"]), float(c.attrib["d"])) for c in (cons if cons is not None else [])]

    names, defaults = [], []
    for force in root.find("Forces"):
        gp = force.find("GlobalParameters")
        if gp is not None:
            for p in gp:
                if p.attrib["name"] not in names:
                    names.append(p.attrib["name"])
                    defaults.append(float(p.attrib["default"]))

    def slot(name):
        return names.index(name)

    bonds, angles, torsions = [], [], []
    nb_atom = np.zeros((A, 3))
    atom_off, exc, exc_off = [], [], []
    exc_pairs = []
    nb_pairs_mixing = None
    sc_pairs = {}
    sc_slots = (-1, -1, -1, -1)
    ext = []

    for force in root.find("Forces"):
        t = force.attrib["type"]
        energy = force.attrib.get("energy", "").replace(" ", "")
        if t == "HarmonicBondForce":
            for b in force.find("Bonds"):
                bonds.append((int(b.attrib["p1"]), int(b.attrib["p2"]), -1, float(b.attrib["k"]), 0.0, float(b.attrib["d"]), 0.0))
        elif t == "CustomBondForce" and energy.startswith("(K/2)*(r-length)^2"):
            pn = [p.attrib["name"] for p in force.find("PerBondParameters")]
            assert pn == ["length1", "K1", "length2", "K2"], pn
            s = slot("lambda_bonds")
            for b in force.find("Bonds"):
                l1, k1, l2, k2 = (float(b.attrib[f"param{i}"]) for i in (1, 2, 3, 4))
                bonds.append((int(b.attrib["p1"]), int(b.attrib["p2"]), s, k1, k2 - k1, l1, l2 - l1))
        elif t == "CustomBondForce" and energy.startswith("U_electrostatics+U_sterics"):
            if len(force.find("Bonds")):
                raise NotImplementedError("unique-atom 1-4 CustomBondForce terms are not lowered yet (0 terms in every vacuum fixture)")
        elif t == "HarmonicAngleForce":
            for a in force.find("Angles"):
                angles.append((int(a.attrib["p1"]), int(a.attrib["p2"]), int(a.attrib["p3"]), -1, float(a.attrib["k"]), 0.0, float(a.attrib["a"]), 0.0))
        elif t == "CustomAngleForce":
            pn = [p.attrib["name"] for p in force.find("PerAngleParameters")]
            assert energy.startswith("(K/2)*(theta-theta0)^2") and pn == ["theta0_1", "K_1", "theta0_2", "K_2"], (energy[:30], pn)
            s = slot("lambda_angles")
            for a in force.find("Angles"):
                t1, k1, t2, k2 = (float(a.attrib[f"param{i}"]) for i in (1, 2, 3, 4))
                angles.append((int(a.attrib["p1"]), int(a.attrib["p2"]), int(a.attrib["p3"]), s, k1, k2 - k1, t1, t2 - t1))
        elif t == "PeriodicTorsionForce":
            for x in force.find("Torsions"):
                n = float(x.attrib["periodicity"])
                assert n == int(n) and 0 <= n <= 6, f"periodicity {n} unsupported"
                torsions.append(tuple(int(x.attrib[f"p{i}"]) for i in (1, 2, 3, 4)) + (-1, int(n), float(x.attrib["k"]), 0.0, float(x.attrib["phase"])))
        elif t == "CustomTorsionForce":
            pn = [p.attrib["name"] for p in force.find("PerTorsionParameters")]
            assert energy.startswith("(1-lambda_torsions)*U1+lambda_torsions*U2") and pn == ["periodicity1", "phase1", "K1", "periodicity2", "phase2", "K2"]
            s = slot("lambda_torsions")
            for x in force.find("Torsions"):
                n1, ph1, k1, n2, ph2, k2 = (float(x.attrib[f"param{i}"]) for i in range(1, 7))
                assert n1 == int(n1) and n2 == int(n2) and 0 <= n1 <= 6 and 0 <= n2 <= 6
                at = tuple(int(x.attrib[f"p{i}"]) for i in (1, 2, 3, 4))
                torsions.append(at + (s, int(n1), k1, -k1, ph1))   # (1-lambda) K1 (1+cos(n1 phi - phase1))
                torsions.append(at + (s, int(n2), 0.0, k2, ph2))   # lambda K2 (1+cos(n2 phi - phase2))
        elif t == "NonbondedForce":
            if force.attrib["method"] != "0":
                raise NotImplementedError("only NoCutoff (vacuum) NonbondedForce is in scope")
            for i, p in enumerate(force.find("Particles")):
                nb_atom[i] = (float(p.attrib["q"]), float(p.attrib["sig"]), float(p.attrib["eps"]))
            po = force.find("ParticleOffsets")
            for o in (po if po is not None else []):
                row = (int(o.attrib["particle"]), slot(o.attrib["parameter"]), float(o.attrib["q"]), float(o.attrib["sig"]), float(o.attrib["eps"]))
                if any(row[2:]):
                    atom_off.append(row)
            for e in force.find("Exceptions"):
                exc_pairs.append((int(e.attrib["p1"]), int(e.attrib["p2"])))
                exc.append((float(e.attrib["q"]), float(e.attrib["sig"]), float(e.attrib["eps"])))
            eo = force.find("ExceptionOffsets")
            for o in (eo if eo is not None else []):
                row = (int(o.attrib["exception"]), slot(o.attrib["parameter"]), float(o.attrib["q"]), float(o.attrib["sig"]), float(o.attrib["eps"]))
                if any(row[2:]):
                    exc_off.append(row)
            nb_pairs_mixing = True
        elif t == "CustomNonbondedForce":
            if force.attrib["method"] != "0":
                raise NotImplementedError("only NoCutoff CustomNonbondedForce is in scope")
            assert energy.startswith("U_sterics;U_sterics=select(step(r-r_LJ)"), energy[:50]
            pn = [p.attrib["name"] for p in force.find("PerParticleParameters")]
            assert pn == ["sigmaA", "epsilonA", "sigmaB", "epsilonB", "unique_old", "unique_new"], pn
            par = np.array([[float(p.attrib[f"param{i}"]) for i in range(1, 7)] for p in force.find("Particles")])
            excl = set()
            for e in force.find("Exclusions"):
                i, j = int(e.attrib["p1"]), int(e.attrib["p2"])
                excl.add((min(i, j), max(i, j)))
            groups = force.find("InteractionGroups")
            inter = set()
            if groups is None or len(groups) == 0:
                inter = {(i, j) for i in range(A) for j in range(i + 1, A)}
            else:
                for g in groups:
                    s1 = [int(p.attrib["index"]) for p in g.find("Set1")]
                    s2 = [int(p.attrib["index"]) for p in g.find("Set2")]
                    inter |= {(min(i, j), max(i, j)) for i in s1 for j in s2 if i != j}
            for (i, j) in sorted(inter - excl):
                epsA, epsB = np.sqrt(par[i, 1] * par[j, 1]), np.sqrt(par[i, 3] * par[j, 3])
                sigA, sigB = 0.5 * (par[i, 0] + par[j, 0]), 0.5 * (par[i, 2] + par[j, 2])
                old, new = max(par[i, 4], par[j, 4]), max(par[i, 5], par[j, 5])
                core = (par[i, 4] + par[j, 4] + par[i, 5] + par[j, 5]) == 0
                # the expression adds core*l_core + new*l_insert + old*l_delete; exactly one class is set in perses hybrids
                assert (1 if core else 0) + int(new) + int(old) == 1, "pair touching both a unique-old and a unique-new atom"
                cls = 0 if core else (1 if new else 2)
                if epsA == 0.0 and epsB == 0.0:
                    continue  # contributes exactly zero for every lambda
                sc_pairs[(i, j)] = (cls, sigA, epsA, sigB, epsB)
            sc_slots = (slot("lambda_sterics_core"), slot("lambda_sterics_insert"), slot("lambda_sterics_delete"), slot("softcore_alpha"))
        elif t == "CustomExternalForce":
            assert energy == "(K/2)*((x-x0)^2+y^2+z^2)+U0", energy
            gn = [p.attrib["name"] for p in force.find("GlobalParameters")]
            pre = [g for g in gn if g.endswith("_K")][0][:-2]
            for p in force.find("Particles"):
                ext.append((int(p.attrib["index"]), slot(pre + "_K"), slot(pre + "_x0"), slot(pre + "_U0")))
        elif t == "CMMotionRemover":
            pass
        else:
            raise NotImplementedError(f"force type {t} ({energy[:40]}) is outside the vacuum hot path")

    # ---- unified pair list
    pairs = {}
    if nb_pairs_mixing:
        exc_index = {(min(i, j), max(i, j)): k for k, (i, j) in enumerate(exc_pairs)}
        off_by_exc = {}
        for row in exc_off:
            off_by_exc.setdefault(row[0], []).append(row)
        off_atoms = {row[0] for row in atom_off}
        for i in range(A):
            for j in range(i + 1, A):
                k = exc_index.get((i, j))
                if k is None:
                    qi_zero = nb_atom[i, 0] == 0 and i not in off_atoms
                    qj_zero = nb_atom[j, 0] == 0 and j not in off_atoms
                    ei_zero = nb_atom[i, 2] == 0 and i not in off_atoms
                    ej_zero = nb_atom[j, 2] == 0 and j not in off_atoms
                    if (qi_zero or qj_zero) and (ei_zero or ej_zero):
                        continue
                    pairs[(i, j)] = -2
                else:
                    if exc[k][0] == 0 and exc[k][2] == 0 and k not in off_by_exc:
                        continue  # a plain exclusion
                    pairs[(i, j)] = k
    keys = sorted(set(pairs) | set(sc_pairs))  # sorted by (i, j): the kernel keeps atom i in registers over a run of pairs
    pair_atoms = [k for k in keys]
    pair_nb = [pairs.get(k, -1) for k in keys]
    pair_sc_class = [sc_pairs[k][0] if k in sc_pairs else -1 for k in keys]
    pair_sc_par = [sc_pairs[k][1:] if k in sc_pairs else (0.0, 0.0, 0.0, 0.0) for k in keys]

    return SystemSpec(
        masses=masses, lambda_names=names, lambda_defaults=np.array(defaults, dtype=np.float64),
        bond_atoms=_i32([b[:2] for b in bonds], 2), bond_slot=_i32([b[2] for b in bonds], 1), bond_par=_f64([b[3:] for b in bonds], 4),
        angle_atoms=_i32([a[:3] for a in angles], 3), angle_slot=_i32([a[3] for a in angles], 1), angle_par=_f64([a[4:] for a in angles], 4),
        torsion_atoms=_i32([x[:4] for x in torsions], 4), torsion_slot=_i32([x[4] for x in torsions], 1),
        torsion_n=_i32([x[5] for x in torsions], 1), torsion_par=_f64([x[6:] for x in torsions], 3),
        nb_atom_par=np.ascontiguousarray(nb_atom), atom_off_idx=_i32([r[:2] for r in atom_off], 2), atom_off_par=_f64([r[2:] for r in atom_off], 3),
        exc_par=_f64(exc, 3), exc_off_idx=_i32([r[:2] for r in exc_off], 2), exc_off_par=_f64([r[2:] for r in exc_off], 3),
        pair_atoms=_i32(pair_atoms, 2), pair_nb=_i32(pair_nb, 1), pair_sc_class=_i32(pair_sc_class, 1), pair_sc_par=_f64(pair_sc_par, 4),
        sc_slots=sc_slots,
        ext_atom=_i32([e[0] for e in ext], 1), ext_slot=_i32([e[1:] for e in ext], 3),
        ext_par=_f64([(0.0, 1.0, 0.0, 1.0, 0.0, 1.0) for _ in ext], 6),
        cons_atoms=_i32([c[:2] for c in cons_rows], 2), cons_dist=_f64([c[2] for c in cons_rows], 1), xml=xml_text)
