"""The NCCL transport of the sharded engine, when the box has at least two GPUs (skipped otherwise):
launches tests/nccl_shard_check.py under torchrun with one rank per GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus() -> int:
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("scenario", ["circle", "lattice", "rings", "late"])
def test_nccl_shards_match_single_gpu_and_oracle(scenario):
    n = _gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    ws = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ws}", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "nccl_shard_check.py"), scenario]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "NCCL-SHARDS-OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
