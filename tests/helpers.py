"""Shared helpers for the test-suite: golden fixtures and small generators."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
VACUUM_SYSTEMS = ["fluorobenzene", "phenol", "aniline", "methylbenzene", "anisole", "1-methylaniline", "ethylbenzene"]


def system_xml(name):
    from oracle.potential import load_system_xml
    return load_system_xml(os.path.join(GOLDEN, "systems", f"{name}.vacuum.xml.gz"))


def system_meta(name):
    with open(os.path.join(GOLDEN, "systems", f"{name}.vacuum.meta.json")) as fh:
        return json.load(fh)


def hybrid_positions(name):
    return np.array(system_meta(name)["hybrid_positions_nm"], dtype=np.float64)


def cache_frames():
    return np.load(os.path.join(GOLDEN, "fluorobenzene_cache.npy"))


def eq0000_frames():
    return np.load(os.path.join(GOLDEN, "tester_eq0000.npy"))


def reference_numbers():
    with open(os.path.join(GOLDEN, "reference_numbers.json")) as fh:
        return json.load(fh)


def generic_layer():
    with open(os.path.join(GOLDEN, "generic_layer.json")) as fh:
        return json.load(fh)


def relative_lambdas(q):
    """coddiwomple/tests/legacy_tests.py:58-76 evaluated at q."""
    return {"lambda_sterics_core": q, "lambda_electrostatics_core": q, "lambda_sterics_insert": 2.0 * q if q < 0.5 else 1.0,
            "lambda_sterics_delete": 0.0 if q < 0.5 else 2.0 * (q - 0.5), "lambda_electrostatics_insert": 0.0 if q < 0.5 else 2.0 * (q - 0.5),
            "lambda_electrostatics_delete": 2.0 * q if q < 0.5 else 1.0, "lambda_bonds": q, "lambda_angles": q, "lambda_torsions": q}


def jittered(x, n, seed=0, scale=0.004):
    rng = np.random.RandomState(seed)
    return (x[None] + rng.normal(0, scale, (n,) + x.shape)).astype(np.float32).astype(np.float64)


