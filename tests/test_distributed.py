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


def n_gpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.gpu
@pytest.mark.parametrize("path,env", [("pipelined", {}), ("p2p", {"PQB_REMAP_SLICE_BITS": "0"}), ("nccl", {"PQB_REMAP_P2P": "0"})])
def test_sharded_state_parity(path, env):
    """every operation of the seam on a sharded state vs the oracle, on as many GPUs as the box has (2, 4 or 8), once per
    remap path: the default (peer-memory exchange kernel, pipelined slice by slice against the neighbouring passes), the same
    kernel without the pipeline, and NCCL send/recv"""
    world = 1
    while world * 2 <= min(n_gpus(), 8):
        world *= 2
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    out = torchrun("dist_check.py", world, 900, PQB_EXPECT=path, **env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "dist_check OK" in out.stdout


@pytest.mark.gpu
def test_broken_remap_is_detected():
    """fault injection: with one peer's sub-block left in place the norm is still 1, and the amplitude comparison must fail"""
    if n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    out = torchrun("dist_check.py", 2, 900, PQB_EXPECT="pipelined", PQB_TEST_BREAK_REMAP="1")
    assert out.returncode != 0 and "AssertionError" in (out.stdout + out.stderr), out.stdout[-2000:] + out.stderr[-2000:]
