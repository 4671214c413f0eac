"""Sharded paths on real GPUs: the cross-frame sweep through eaof/sweep.py on one GPU against the
oracle, and — when the box has at least two GPUs — frame-sharded extraction + NCCL all-gather + partitioned sweep under
torchrun against the single-GPU result (tools/multi_gpu_check.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, port):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)


def test_sweep_single_gpu_against_oracle():
    out = _run(1, 29541)
    assert "MULTI_GPU_CHECK world=1" in out.stdout and " OK" in out.stdout, out.stdout + out.stderr[-3000:]


def test_sweep_two_gpus_nccl_allgather():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    out = _run(2, 29542)
    assert "MULTI_GPU_CHECK world=2" in out.stdout and " OK" in out.stdout, out.stdout + out.stderr[-3000:]
