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


@pytest.mark.skipif(n_gpus() < 2, reason="needs at least two GPUs")
@pytest.mark.parametrize("world,p2p", [(2, "0"), (2, "1"), (4, "1")])
def test_single_process_multi_gpu_handle_reproduces_the_golden_loop(world, p2p, monkeypatch):
    """esimp_multi_* (one process, one host thread per GPU inside the library -- how a single Julia session reaches N GPUs): BASELINE
    config 1 with the multigrid-preconditioned solve against the golden fixture and against a one-GPU context; same bars as the
    one-process-per-GPU path (compliance 1e-6, density 1e-3).  p2p = 1: node-plane halos over peer memory (same address space: plain
    peer pointers instead of CUDA IPC)."""
    if n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import numpy as np

    sys.path[:0] = [str(ROOT), str(ROOT / "easysimp.jl_b200")]
    import easysimp_b200 as es
    from easysimp_b200.lib import MultiContext
    from oracle import simp_oracle as o

    monkeypatch.setenv("ESIMP_P2P_HALO", p2p)
    gz = np.load(ROOT / "tests" / "golden" / "c1_cantilever_60x20x4.npz")
    g, dofs, f, pres = o.cantilever_problem((60, 20, 4))
    prm = dict(filter_radius=1.5 * o.estimate_element_size(g), pcg_rtol=1e-10, precond=1)

    def run(ctx, n_it):
        ctx.set_params(**prm)
        ctx.set_grid_structured(60, 20, 4, 1.0, 1.0, 1.0, dofs)
        ctx.set_dirichlet(pres)
        ctx.set_load(f)
        ctx.filter_build_structured()
        ctx.init_design(None)
        st = [ctx.iterate() for _ in range(n_it)]
        return st

    m = MultiContext(n_gpus=world)
    assert m.n_gpus == world
    st = run(m, 50)
    energy = np.array([s.energy for s in st])
    assert np.max(np.abs(energy - gz["energy"]) / gz["energy"]) < 1e-6
    e_fin, _, _ = m.final_analysis()
    assert abs(e_fin - gz["final_energy"][0]) / gz["final_energy"][0] < 1e-6
    rho_phys, rho, u = m.get_physical_density(), m.get_design(), m.get_displacement()
    assert np.abs(rho_phys - gz["final_density"]).max() < 1e-3
    # a host-carried design goes back in slab by slab and the loop continues from it
    m.set_design(rho)
    s_next = m.iterate()
    m.close()
    one = es.Context(0)
    st1 = run(one, 50)
    e1, _, _ = one.final_analysis()
    assert abs(e_fin - e1) / e1 < 1e-9
    u1 = one.get_displacement()
    assert np.abs(u - u1).max() / np.abs(u1).max() < 1e-6 and np.abs(rho - one.get_design()).max() < 1e-6
    one.set_design(one.get_design())
    assert abs(s_next.energy - one.iterate().energy) / s_next.energy < 1e-8
    one.close()
