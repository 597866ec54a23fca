import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
os.environ.setdefault("DNET_TRANSPORT_WIRE_DTYPE", "bf16")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_lib():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dnet_b200 import _cabi

    _cabi.init(0)
    return _cabi.load()
