"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm (`--impl reference`) prints ONE JSON line with
the agreed keys, times the compiled CPU port on every host core even when OMP_NUM_THREADS=1 is exported (torchrun does that to its ranks), and
non-zero ranks stay silent; the frozen algorithmic constants reproduce SURVEY.md §8(d)'s figures."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env):
    env = dict(os.environ, **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1"], env=env,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_line():
    out = _run({"OMP_NUM_THREADS": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle-Langevin-steps/s" and d["unit"] == "particle-steps/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and "sample" in cb
    n_cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    assert cb["cores"] == n_cores                      # not the 1 thread OMP_NUM_THREADS asked for
    assert abs(d["free_energy_kT"] - 1.33) < 0.2        # the reference's own band for this system (tests/legacy_tests.py:90-94)


def test_other_ranks_of_the_reference_arm_print_nothing():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_frozen_algorithmic_constants():
    sys.path.insert(0, ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from coddiwomple_b200 import testsystems as ts
    from coddiwomple_b200.system import spec_from_xml
    from tests.helpers import system_xml
    assert bench.algorithmic_flops(spec_from_xml(system_xml("fluorobenzene"))) == {"f_eval": 8707, "per_iteration": 18734}
    assert bench.algorithmic_flops(spec_from_xml(ts.alanine_dipeptide_shaped_xml()[0]))["per_iteration"] == 40498
