import importlib
import os
import sys

import numpy as np
import pytest

# before CUDA initialises: one hardware queue per stream for the in-process rank groups (W contexts x 2 streams on one device),
# and a short leash on the peer-memory exchange's device-side waits (a lost peer becomes an error, never a hang)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("DPRT_P2P_TIMEOUT_MS", "8000")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def dprt():
    return importlib.import_module("pg2024-data-parallel-ray-tracing_b200")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_required():
    if not has_gpu():
        pytest.fail("this test is marked gpu but no CUDA device is visible")
    return True
