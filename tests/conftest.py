"""pytest configuration: marker registration and import paths.

`-m "not gpu"` runs on the CPU-only build container (oracle vs golden vectors, host logic, C-ABI symbols);
`-m gpu` needs a B200 and is where the CUDA path is compared with the oracle.
"""
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "omnilingual-asr_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "slow: several seconds of CPU work")


@pytest.fixture(scope="session")
def device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("CUDA not available")
    return torch.device("cuda", 0)
