"""Race check of the team algorithm without a GPU (compute-sanitizer's racecheck needs one): the host harness is built
with ThreadSanitizer and the T lanes of a replica run as T host threads over the same "shared memory" columns; a missing
CDW_TEAM_SYNC in coddiwomple_b200/csrc/cdw_lane.cuh is reported as a data race.  (This is how the missing barrier between
the constraint-vector snapshot and the atom moves of the Euler-Maruyama op was found.)"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HARNESS = os.path.join(HERE, "host_harness")
TSAN_SO = os.path.join(HARNESS, "libcdw_host_harness_tsan.so")


def _libtsan():
    try:
        path = subprocess.check_output(["gcc", "-print-file-name=libtsan.so"], text=True).strip()
    except (OSError, subprocess.CalledProcessError):
        return None
    return path if os.path.isabs(path) and os.path.exists(path) else None


@pytest.fixture(scope="module")
def tsan_build():
    lib = _libtsan()
    if lib is None:
        pytest.skip("libtsan is not available")
    src = os.path.join(HARNESS, "harness.cpp")
    deps = [src] + [os.path.join(ROOT, "coddiwomple_b200", "csrc", f) for f in ("cdw_math.cuh", "cdw_lane.cuh", "cdw_pack.h")]
    if not os.path.exists(TSAN_SO) or any(os.path.getmtime(d) > os.path.getmtime(TSAN_SO) for d in deps):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++20", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-fsanitize=thread", "-o", TSAN_SO, src],
                              cwd=HARNESS)
    return lib


@pytest.mark.parametrize("system,T,mode", [("fluorobenzene", 4, "baoab"), ("fluorobenzene", 4, "cached"), ("fluorobenzene", 4, "ula"),
                                           ("ala", 4, "cached"), ("ala", 8, "baoab"), ("fluorobenzene", 4, "lookahead"), ("ala", 4, "lookahead"), ("fluorobenzene", 4, "twisted")])
def test_team_algorithm_has_no_data_race(tsan_build, system, T, mode):
    env = dict(os.environ, LD_PRELOAD=tsan_build, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0", CDW_HH_TSAN_SO=TSAN_SO)
    env.pop("CDW_HH_TEAM", None)
    out = subprocess.run([sys.executable, os.path.join(HARNESS, "racecheck_driver.py"), system, str(T), mode], env=env, capture_output=True, text=True,
                         timeout=900)
    assert "racecheck-done" in out.stdout, out.stderr[-2000:]
    races = [ln for ln in out.stderr.splitlines() if "ThreadSanitizer: data race" in ln]
    assert not races, "\n".join(out.stderr.splitlines()[:60])
