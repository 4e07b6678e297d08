#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ from the read-only reference checkout.

Run in the build container only (needs /root/reference); the GPU box never runs this. Nothing here is
imported by the product package.

What is extracted (all cited relative to /root/reference):

* systems/<name>.vacuum.xml.gz   the `_hybrid_system` OpenMM-7.4 System XML embedded in
  coddiwomple/data/perses_data/benzene_<name>.vacuum.factory.pkl (the longest <System>…</System> block), plus
  systems/<name>.vacuum.meta.json with `_hybrid_positions` (nm) and `_atom_classes` read through a stub Unpickler
  (no perses/simtk needed).
* fluorobenzene_cache.npy        the 151 λ=0 equilibrium frames of
  coddiwomple/data/perses_data/fluorobenzene_zero_endstate_cache/vacuum.cache.pdb as float32 nm
  (float32(Å)/float32(10), which is what mdtraj hands the reference at openmm/coddiwomple.py:286,304).
* tester_eq0000.npy              the 10 frames of troubleshooting/tester/eq.0000.pdb, same convention
  (input of the executed notebook troubleshooting/openmm_smc.ipynb).
* reference_numbers.json         numbers the reference itself logged (notebook JSON line numbers given).
* generic_layer.json             outputs of the reference's importable generic layer
  (coddiwomple.{utils,resamplers,particles,distribution_factories}) on seeded inputs, produced by importing it
  from /root/reference.
"""
import gzip
import json
import os
import pickle
import re
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["fluorobenzene", "phenol", "aniline", "methylbenzene", "anisole", "1-methylaniline", "ethylbenzene"]


class _Stub:
    def __init__(self, *a, **k):
        self._args = a

    def __setstate__(self, st):
        if isinstance(st, dict):
            self.__dict__.update(st)
        else:
            self._state = st

    def __call__(self, *a, **k):
        return _Stub()


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] in ("numpy", "builtins", "collections", "copyreg", "_codecs"):
            return super().find_class(module, name)
        return type(name, (_Stub,), {})


def read_pdb_frames(path):
    frames, cur = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("ATOM") or line.startswith("HETATM"):
                cur.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
            elif line.startswith("ENDMDL"):
                frames.append(cur)
                cur = []
    if cur:
        frames.append(cur)
    ang = np.asarray(frames, dtype=np.float32)
    return (ang / np.float32(10.0)).astype(np.float32)


def extract_systems():
    os.makedirs(os.path.join(HERE, "systems"), exist_ok=True)
    for name in NAMES:
        pkl = f"{REF}/coddiwomple/data/perses_data/benzene_{name}.vacuum.factory.pkl"
        raw = open(pkl, "rb").read()
        blocks = re.findall(r"<System .*?</System>", raw.decode("latin-1"), flags=re.S)
        xml = max(blocks, key=len)
        with gzip.GzipFile(os.path.join(HERE, "systems", f"{name}.vacuum.xml.gz"), "wb", mtime=0) as fh:
            fh.write(xml.encode("ascii"))
        fac = _StubUnpickler(open(pkl, "rb")).load()
        meta = {
            "source": f"coddiwomple/data/perses_data/benzene_{name}.vacuum.factory.pkl",
            "hybrid_positions_nm": np.asarray(fac._hybrid_positions._value, dtype=np.float64).tolist(),
            "atom_classes": {k: sorted(int(i) for i in v) for k, v in fac._atom_classes.items()},
            "softcore_LJ_v2": bool(fac._softcore_LJ_v2),
            "interpolate_14s": bool(fac._interpolate_14s),
        }
        with open(os.path.join(HERE, "systems", f"{name}.vacuum.meta.json"), "w") as fh:
            json.dump(meta, fh, indent=1)
        print("system", name, len(xml), "bytes of XML;", len(meta["hybrid_positions_nm"]), "atoms")


def extract_frames():
    cache = read_pdb_frames(f"{REF}/coddiwomple/data/perses_data/fluorobenzene_zero_endstate_cache/vacuum.cache.pdb")
    np.save(os.path.join(HERE, "fluorobenzene_cache.npy"), cache)
    tester = read_pdb_frames(f"{REF}/troubleshooting/tester/eq.0000.pdb")
    np.save(os.path.join(HERE, "tester_eq0000.npy"), tester)
    print("frames", cache.shape, tester.shape)


def reference_numbers():
    """Numbers logged by the reference's own executed notebook troubleshooting/openmm_smc.ipynb."""
    nb = open(f"{REF}/troubleshooting/openmm_smc.ipynb").read().split("\n")

    def grab(lineno, pattern):
        m = re.search(pattern, nb[lineno - 1])
        assert m, (lineno, nb[lineno - 1][:200])
        return m.group(1)

    nums = {
        "source": "troubleshooting/openmm_smc.ipynb (JSON line numbers, 1-based)",
        "u_lambda0": {
            # (value, JSON line, frame index of troubleshooting/tester/eq.0000.pdb identified in SURVEY.md §8c)
            "frame2": [float(grab(260, r"state reduced potential \(([-0-9.e]+)\)")), 260],
            "frame0": [float(grab(264, r"state reduced potential \(([-0-9.e]+)\)")), 264],
            "frame1": [float(grab(268, r"state reduced potential \(([-0-9.e]+)\)")), 268],
        },
        "first_mean_cumulative_work": [float(grab(339, r"mean cumulative work: ([-0-9.e]+)")), 339],
        "particle_frame_order": [2, 0, 1],
    }
    # the whole trace of "mean cumulative work" values and resampled indices (G5, statistical)
    trace, draws = [], []
    for i, line in enumerate(nb):
        m = re.search(r"mean cumulative work: ([-0-9.e]+)", line)
        if m:
            trace.append([float(m.group(1)), i + 1])
        m = re.search(r"resampled_indices : \[([0-9 ]+)\]", line)
        if m:
            draws.append([[int(t) for t in m.group(1).split()], i + 1])
    nums["mean_cumulative_work_trace"] = trace
    nums["resampled_indices_trace"] = draws
    nums["legacy_test_band"] = {"center": 1.33, "tolerance": 0.2, "source": "coddiwomple/tests/legacy_tests.py:90-94"}
    with open(os.path.join(HERE, "reference_numbers.json"), "w") as fh:
        json.dump(nums, fh, indent=1)
    print("reference numbers:", nums["u_lambda0"], nums["first_mean_cumulative_work"], len(trace), "trace points")


def generic_layer():
    """Drive the reference's importable generic layer on seeded inputs and record what it returns."""
    sys.path.insert(0, REF)
    import logging

    logging.disable(logging.CRITICAL)
    from coddiwomple.particles import Particle
    from coddiwomple.resamplers import MultinomialResampler, Resampler
    from coddiwomple.utils import nESS, normalized_weights, vanilla_floor_threshold

    out = {"source": "coddiwomple/{utils,resamplers,particles}.py imported from /root/reference", "cases": []}
    for case, (P, scale, seed) in enumerate([(3, 0.3, 1), (16, 1.0, 2), (257, 3.0, 3), (1000, 0.1, 4), (1000, 2.0, 5), (4096, 5.0, 6)]):
        rng = np.random.RandomState(seed)
        cum = rng.normal(0.0, scale, P)
        inc = rng.normal(0.5, scale, P)
        aux = rng.normal(0.0, 1.0, P)
        particles = [Particle(index=i) for i in range(P)]
        for p, c, a in zip(particles, cum, aux):
            p._cumulative_work = float(c)
            p._auxiliary_work = float(a)
            p.update_state(("state", p.ancestry[0]))
        ness = float(nESS(particles, incremental_works=inc.copy()))
        do = bool(vanilla_floor_threshold(ness))
        w = normalized_weights(cum + inc)
        # the reference's multinomial draw: np.random.choice under the GLOBAL numpy seed (resamplers.py:271)
        np.random.seed(1000 + seed)
        res = MultinomialResampler()
        res.resample(particles, inc.copy(), observable=nESS, threshold=vanilla_floor_threshold)
        # the uniforms np.random.choice consumed (numpy: cdf.searchsorted(random_sample(size), side='right'))
        np.random.seed(1000 + seed)
        uniforms = np.random.random_sample(P)
        # forced multinomial draw (observable=None ⇒ always resample, resamplers.py:90-92)
        particles2 = [Particle(index=i) for i in range(P)]
        for p, c, a in zip(particles2, cum, aux):
            p._cumulative_work = float(c)
            p._auxiliary_work = float(a)
            p.update_state(("state", p.ancestry[0]))
        np.random.seed(1000 + seed)
        MultinomialResampler().resample(particles2, inc.copy())
        out["cases"].append({
            "P": P, "seed": seed, "scale": scale,
            "cumulative_works_in": cum.tolist(), "incremental_works": inc.tolist(), "auxiliary_works_in": aux.tolist(),
            "nESS": ness, "threshold_says_resample": do, "normalized_weights": w.tolist(),
            "resampling_logger": res.resampling_logger,
            "after_thresholded": {
                "cumulative_work": [p.cumulative_work for p in particles],
                "auxiliary_work": [p.auxiliary_work for p in particles],
                "ancestry_last": [int(p.ancestry[-1]) for p in particles],
                "state_label": [int(p.state[1]) for p in particles],
            },
            "uniforms": uniforms.tolist(),
            "after_forced": {
                "cumulative_work": [p.cumulative_work for p in particles2],
                "auxiliary_work": [p.auxiliary_work for p in particles2],
                "ancestry_last": [int(p.ancestry[-1]) for p in particles2],
                "state_label": [int(p.state[1]) for p in particles2],
            },
        })
        print("generic case", case, "P", P, "nESS", ness, "resample", do)

    # the SIS base class (identity _resample, resamplers.py:202-216) on one case
    P = 8
    rng = np.random.RandomState(11)
    inc = rng.normal(0, 1, P)
    particles = [Particle(index=i) for i in range(P)]
    for p in particles:
        p.update_state(("state", p.ancestry[0]))
    Resampler().resample(particles, inc.copy())
    out["sis_case"] = {"incremental_works": inc.tolist(), "cumulative_work": [p.cumulative_work for p in particles],
                       "ancestry_last": [int(p.ancestry[-1]) for p in particles]}
    with open(os.path.join(HERE, "generic_layer.json"), "w") as fh:
        json.dump(out, fh)


if __name__ == "__main__":
    extract_systems()
    extract_frames()
    reference_numbers()
    generic_layer()
