"""Test helpers: bind the package to the real CUDA library or to the emulated build."""

import os

import pytest


def bind_emulation():
    """CPU tests: run the same C-ABI code compiled against the host emulation (tests/emu)."""
    from tests.emu.build_emu import build

    from tensorsat_b200 import _native

    lib = build()
    if _native._device != "cpu" or _native._lib is None:
        _native._testing_bind(lib, "cpu")


def bind_cuda():
    from tensorsat_b200 import _native

    if not os.path.exists(_native.LIB_PATH):
        pytest.fail("libtsb.so is missing on a GPU box: run python -m tensorsat_b200.build")
    if _native._device != "cuda":
        _native._testing_bind(_native.LIB_PATH, "cuda")
