"""Multi-process slab partition on real GPUs (needs >= 2 devices; skipped on a 1-GPU box): torchrun launches
tests/dist/dist_check.py, which exchanges the CUDA IPC handles through torch.distributed and compares FormFunction /
MatMult_Elliptic of every rank with the oracle's single-domain result (1e-12 max-norm relative)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_multi_process(world):
    if _ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + world), os.path.join(ROOT, "tests", "dist", "dist_check.py"), "32", "64"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith('{"check"')]
    assert out.returncode == 0, out.stderr[-2000:]
    assert len(lines) == 3 and all(l["ok"] and l["ranks"] == world for l in lines)  # slab at 32 and 64, slab KSP at 32


@pytest.mark.parametrize("world", [2, 8])
def test_stokes_slab_multi_process(world):
    if _ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29720 + world), os.path.join(ROOT, "tests", "dist", "dist_stokes.py"), "16", "0"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith('{"check"')]
    assert out.returncode == 0, out.stderr[-2000:]
    assert len(lines) == 1 and lines[0]["ok"] and lines[0]["ranks"] == world


@pytest.mark.parametrize("world", [2, 8])
def test_slab_saddle_solve_multi_process(world):
    """Config 5 as a SOLVE on the slab partition: outer FGMRES + block-LU saddle PC with slab inner solvers (tests/dist/dist_saddle.py)."""
    if _ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29740 + world), os.path.join(ROOT, "tests", "dist", "dist_saddle.py"), "16" if world == 8 else "32"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith('{"check"')]
    assert out.returncode == 0, out.stdout[-1000:] + out.stderr[-2000:]
    assert len(lines) == 1 and lines[0]["ok"] and lines[0]["ranks"] == world
