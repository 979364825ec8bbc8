"""Multi-GPU parity (needs >= 2 visible GPUs; skipped otherwise): the sharded stepper in both
decompositions reproduces the single-GPU trajectory."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [16384, 9000])
def test_sharded_stepper_matches_single_gpu(n):
    import torch

    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if ng < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29641", os.path.join(ROOT, "tests", "mgpu_worker.py"), str(n)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_slab_stepper_matches_single_gpu():
    import torch

    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if ng < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29643", os.path.join(ROOT, "tests", "mgpu_worker.py"), "slab", "16"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_c_abi_group_across_processes():
    """nbx_group_* with CUDA IPC handles between torchrun ranks: slabs bit-identical to one GPU, pair-sharded gravity,
    target-block water, Langevin EM -- no collective library on the data path."""
    import torch

    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if ng < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29645", os.path.join(ROOT, "tests", "mgpu_worker.py"), "group"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_nbx_create_multi_over_distinct_devices():
    """One handle, one host thread, several GPUs (no torch.distributed)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "mgpu_worker.py"), "multi"], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    assert "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
