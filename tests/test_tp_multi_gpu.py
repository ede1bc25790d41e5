"""Tensor parallelism over NCCL on >= 2 GPUs of one box (BASELINE config 4): launched as a torchrun subprocess."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("world,mode", [(2, "nccl"), (2, "fused")])
def test_tp_matches_single_gpu(world, mode):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(ROOT / "scripts" / "tp_check.py"),
                        "wide2l", mode],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "TP OK" in r.stdout
