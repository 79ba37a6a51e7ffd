"""rxc_mgpu_* on real GPUs: every rank's raster kernel writes into rank 0's delivery buffer (peer mapping over NVLink,
or the grouped-NCCL fallback) and rank 0 compares with its own single-GPU render bit for bit (tests/mgpu_worker.py).
The one-GPU box runs the world-1 leg; the multi-rank legs need `gpurun --gpus N`."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "mgpu_worker.py")


def _gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _run(world, extra_env=None, timeout=600):
    env = dict(os.environ)
    env.update(extra_env or {})
    if world == 1:
        cmd = [sys.executable, WORKER]
    else:
        port = 29600 + (os.getpid() % 1500)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), WORKER]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("MGPU_WORKER_OK")]
    assert line, r.stdout[-3000:] + r.stderr[-3000:]
    return line[-1]


@pytest.mark.gpu
def test_mgpu_world1_delivery_buffer_matches_direct_render():
    assert "mode=local world=1" in _run(1)


@pytest.mark.gpu
def test_mgpu_world2_peer_writes_are_bit_identical():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    assert "mode=peer world=2" in _run(2)


@pytest.mark.gpu
def test_mgpu_world2_nccl_fallback_is_bit_identical():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    assert "mode=nccl world=2" in _run(2, {"RXC_MGPU_FORCE_NCCL": "1"})


@pytest.mark.gpu
def test_mgpu_all_gpus_of_the_box():
    n = _gpus()
    if n < 4:
        pytest.skip("needs 4+ GPUs")
    assert f"mode=peer world={n}" in _run(n)
