"""-m gpu tests of the x-slab multi-GPU path, run by pytest on any box with at least two GPUs (skipped on one GPU):
tests/mgpu_check.py (operator / filter stage calls against the oracle and the golden loops with the Jacobi and the
multigrid preconditioner, one process per GPU under torch.distributed.run) and tests/mgpu_post_check.py (body force,
stress fields, loops with acceleration_data; plain processes).  Same bars as the 1-GPU tests: compliance 1e-6,
final density 1e-3."""
import socket
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def n_gpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(n_gpus() < 2, reason="needs at least two GPUs")
@pytest.mark.parametrize("world", [2, 4])
def test_slab_parity_under_torchrun(world):
    if n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(free_port()), str(ROOT / "tests" / "mgpu_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "[mgpu] ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.skipif(n_gpus() < 2, reason="needs at least two GPUs")
def test_post_load_rows_on_two_gpus():
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "mgpu_post_check.py"), "2"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "[mgpu-post] ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
