"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol of include/tsb.h."""

import os
import re

from tests.conftest import ROOT


def _declared():
    with open(os.path.join(ROOT, "include", "tsb.h")) as f:
        src = f.read()
    return sorted(set(re.findall(r"^(?:int|const char\*)\s+(tsb_[a-z0-9_]+)\s*\(", src, flags=re.M)))


def test_header_and_binding_agree():
    from tensorsat_b200 import _native

    assert _declared() == sorted(_native.EXPORTS)


def test_library_builds_loads_and_exports_every_symbol():
    from tensorsat_b200 import _native, build

    lib = build.build()
    handle = _native.load_library(lib)
    for name in _declared():
        assert hasattr(handle, name), name
    assert handle.tsb_version() == 100


def test_sass_targets_sm100a():
    import shutil
    import subprocess

    from tensorsat_b200 import build

    if shutil.which("cuobjdump") is None:
        return
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_missing_library_fails_loudly(tmp_path):
    import pytest

    from tensorsat_b200 import _native

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _native.load_library(str(tmp_path / "libtsb.so"))
