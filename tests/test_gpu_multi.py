"""The sharded anneal on every GPU of the box against the single-GPU anneal, bit for bit: tests/multi_gpu_check.py under torchrun, collected by
pytest so that it runs wherever `-m gpu` finds at least two devices (skipped on a one-GPU box; the same check runs inside every multi-GPU
bench.py invocation as `parity_probe`)."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_anneal_equals_single_gpu_anneal_bitwise():
    if not torch.cuda.is_available():
        pytest.fail("the -m gpu tests need a CUDA device: there is no CPU fallback to test")
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU: nothing to shard over")
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    env = dict(os.environ, CDW_CHECK_GRAPH="1", CDW_WATCHDOG_S="300")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                          "--master-port", "29541", os.path.join(ROOT, "tests", "multi_gpu_check.py")], capture_output=True, text=True, timeout=600, env=env)
    lines = [l for l in out.stdout.splitlines() if "[multi_gpu_check]" in l]
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert len(lines) >= 7 and all("True" in l for l in lines), "\n".join(lines)
