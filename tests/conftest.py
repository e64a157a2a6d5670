import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """A plain ``pytest tests`` on a box without CUDA skips the ``gpu`` suites instead of failing; with a GPU present
    they run and fail hard if libtsb.so is missing (tests/backend.py).  ``-m "not gpu"`` is the CPU suite."""
    try:
        import torch

        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_ops():
    return load_golden("ops.json")


@pytest.fixture(scope="session")
def golden_instances():
    return load_golden("instances.json")


@pytest.fixture(scope="session")
def golden_networks():
    return load_golden("networks.json")
