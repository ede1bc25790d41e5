"""Tensor parallelism over NCCL on >= 2 GPUs of one box (BASELINE config 4): launched as a torchrun subprocess."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("world,mode,reps,overlap", [(2, "nccl", 1, "on"), (2, "fused", 1, "on"), (2, "fused", 3, "on"),
                                                     (2, "fused", 3, "off"), (4, "fused", 3, "on"), (8, "fused", 3, "on")])
def test_tp_matches_oracle_and_single_gpu(world, mode, reps, overlap):
    """reps = 1: one fused peer-memory kernel per reduction (three windows); reps = 3: nine ragged windows, the
    half-batch pipeline (the tail of one half-batch's reduction beside the other's compute); OASR_TP_OVERLAP=off runs the
    same nine windows without the pipeline.  All against the oracle's tensor-parallel rounding points and the unsplit
    engine (scripts/tp_check.py)."""
    import os
    env = dict(os.environ, OASR_TP_OVERLAP=overlap)
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(ROOT / "scripts" / "tp_check.py"),
                        "wide2l", mode, str(reps)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "TP OK" in r.stdout
