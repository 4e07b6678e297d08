"""Host logic: System XML -> term tables, and the C ABI library's exported surface.  CPU only (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from coddiwomple_b200 import _build, _lib, testsystems as ts
from coddiwomple_b200.system import spec_from_xml
from tests.helpers import ROOT, VACUUM_SYSTEMS, system_xml

# atoms, constraints, bonds, angles, torsions, exceptions per vacuum fixture
EXPECTED = {"fluorobenzene": (13, 6, 7, 20, 35), "phenol": (14, 7, 7, 21, 37), "aniline": (15, 8, 7, 23, 40), "methylbenzene": (16, 9, 7, 26, 35),
            "anisole": (17, 9, 8, 27, 40), "1-methylaniline": (18, 10, 8, 29, 43), "ethylbenzene": (19, 11, 8, 32, 46)}


@pytest.mark.parametrize("name", VACUUM_SYSTEMS)
def test_term_counts(name):
    c = spec_from_xml(system_xml(name)).counts()
    assert (c["atoms"], c["constraints"], c["bonds"], c["angles"], c["torsions"]) == EXPECTED[name]
    assert c["n_lambda"] == 10  # nine lambdas + softcore_alpha


def test_fluorobenzene_pairs():
    s = spec_from_xml(system_xml("fluorobenzene"))
    # 19 non-excepted pairs + 25 non-zero 1-4 exceptions; plain exclusions are dropped
    assert len(s.pair_nb) == 44 and int((s.pair_nb == -2).sum()) == 19 and int((s.pair_nb >= 0).sum()) == 25
    assert int((s.pair_sc_class >= 0).sum()) == 19 and set(s.pair_sc_class[s.pair_sc_class >= 0]) == {0, 1, 2}
    assert s.lambda_vector({"lambda_bonds": 0.5})[s.lambda_names.index("lambda_bonds")] == 0.5
    assert s.lambda_vector()[s.lambda_names.index("softcore_alpha")] == 0.85
    with pytest.raises(KeyError):
        s.lambda_vector({"lambda_nonexistent": 1.0})


def test_synthetic_alanine_dipeptide_shape():
    xml, x = ts.alanine_dipeptide_shaped_xml()
    c = spec_from_xml(xml).counts()
    assert (c["atoms"], c["constraints"], c["angles"]) == (22, 12, 36) and c["bonds"] == 9 and c["torsions"] == 41 + 6
    assert c["pairs"] == 231 - 21 - 36  # every pair that is not a 1-2 / 1-3 exclusion contributes
    xml2, x2 = ts.alanine_dipeptide_shaped_xml()
    assert xml == xml2 and np.array_equal(x, x2)  # seeded


def test_unsupported_forces_are_rejected_loudly():
    xml = system_xml("fluorobenzene").replace('method="0" nx', 'method="4" nx')
    with pytest.raises(NotImplementedError):
        spec_from_xml(xml)


def test_library_builds_and_exports_every_declared_symbol():
    """The C-ABI library loads here (no GPU) and exports exactly what include/cdw.h declares."""
    _build.build()
    header = open(os.path.join(ROOT, "include", "cdw.h")).read()
    declared = set(re.findall(r"CDW_API\s+[\w\s\*]+?\b(cdw_\w+)\s*\(", header))
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().cdw_version() == 1
    assert _lib.load().cdw_weight_workspace_bytes(10 ** 6) > 8 * (10 ** 6 // 2048)


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(_lib.LangevinParams) == 72   # 4 doubles, 2 x 32 bit, the splitting word, the trial-count, twist and dense-twist pointers
    assert _lib.LangevinParams.dense_twist.offset == 64 and ctypes.sizeof(_lib.DenseTwist) == 64
    assert _lib.LangevinParams.splitting.offset == 40 and _lib.LangevinParams.trial_counts.offset == 48 and _lib.LangevinParams.twist.offset == 56
    # cdw_system_desc: count the fields of the C struct
    header = open(os.path.join(ROOT, "include", "cdw.h")).read()
    body = header[header.index("typedef struct cdw_system_desc {"):header.index("} cdw_system_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    n_fields = sum(len(stmt.split(",")) for stmt in body.split(";") if re.search(r"\b(int32_t|double)\b", stmt))
    assert n_fields == len(_lib.SystemDesc._fields_)


def test_no_cpu_fallback_without_cuda():
    """On a box without a GPU the product refuses to run instead of falling back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from coddiwomple_b200.engine import DeviceSystem
    with pytest.raises(_lib.CdwError):
        DeviceSystem(ts.harmonic_oscillator_xml())


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "coddiwomple_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_entry_points_reject_bad_arguments_without_touching_the_gpu():
    """Error behaviour of the C ABI (include/cdw.h): a negative return code and a thread-local message, never an abort;
    argument validation comes before any CUDA call, so it is testable on a box without a GPU."""
    lib = _lib.load()
    prm = _lib.LangevinParams(0.002, 1.0, 2.494, 1e-6, 1, 0)
    none = None
    calls = {
        "cdw_reduced_potential": (none, none, 4, none, 0.4, none, none),
        "cdw_smc_propagate": (none,) * 5 + (4,) + (none,) * 5 + (ctypes.byref(prm), 0, 0, 0) + (none,) * 11,
        "cdw_smc_propagate_cached": (none,) * 5 + (4,) + (none,) * 5 + (ctypes.byref(prm), 0, 0, 0) + (none,) * 12,
        "cdw_smc_lookahead": (none,) * 7 + (4,) + (none,) * 4 + (ctypes.byref(prm), 0, 0, 0) + (none,) * 5,
        "cdw_smc_lookahead_sharded": (none,) * 5 + (2, 2) + (none,) * 3 + (4,) + (none,) * 3 + (ctypes.byref(prm), 0, 0, 0) + (none,) * 5,
        "cdw_smc_weigh_resample": (0, none, none, 4, 0.5, none, 0, 0) + (none,) * 8 + (0, none),
        "cdw_peer_barrier": (none, none, 0, 2, none, none, none, none, none),
        "cdw_record_lineage": (none,) * 5 + (4, none),
        "cdw_smc_resample": (0, none, 4, 0.5, none, 0, 0) + (none,) * 7 + (0, none),
        "cdw_weight_stats": (none, 4, 0.5, none, none, none, 0, none),
        "cdw_resample_indices": (0, none, none, 4, none, none, none, none, none, none, 0, none),
        "cdw_gather_particles": (none,) * 9 + (4, 3, none),
        "cdw_system_register_protocol": (none, none, 3, 2.494, none),
        "cdw_system_create": (none, none),
    }
    for name, args in calls.items():
        rc = getattr(lib, name)(*args)
        assert rc < 0, name
        assert lib.cdw_last_error(), name
    # well-formed but empty batches are fine and do nothing (n == 0 returns before the null checks)
    assert lib.cdw_smc_propagate(*((none,) * 5 + (0,) + (none,) * 5 + (ctypes.byref(prm), 0, 0, 0) + (none,) * 11)) < 0   # still needs sys
    assert lib.cdw_system_destroy(None) == 0
    with pytest.raises(_lib.CdwError):
        _lib.check(lib.cdw_weight_stats(none, 4, 0.5, none, none, none, 0, none), "cdw_weight_stats")
