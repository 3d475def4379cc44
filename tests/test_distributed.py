"""N > 1 path: gloo world_size-2 plumbing test on the CPU, NCCL sharded-state parity on >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def torchrun(script, nproc, timeout, **extra_env):
    port = free_port()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", **extra_env)
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
                           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", script)],
                          capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_world2_plumbing_gloo_cpu():
    out = torchrun("dist_plumbing_cpu.py", 2, 300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "plumbing OK" in out.stdout


@pytest.mark.gpu
def test_sharded_state_parity_two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    out = torchrun("dist_check.py", 2, 600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "dist_check OK" in out.stdout


@pytest.mark.gpu
def test_sharded_state_parity_two_gpus_peer_memory_exchange():
    """same check with the opt-in peer-memory remap path (exported VMM handles, one swap kernel over NVLink)"""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    out = torchrun("dist_check.py", 2, 600, PQB_REMAP_P2P="1")
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "dist_check OK" in out.stdout and "'p2p_remaps': 0" not in out.stdout
