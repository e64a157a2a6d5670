"""Import the real reference (``/root/reference``) under Python 3.12 (test infrastructure).

The reference needs Python >= 3.13 only for ``TypeVar(..., default=...)`` at three sites
(tensorsat/_utils/meta.py:134,138; diagrams.py:1034) and a handful of pure-Python
third-party modules absent from this image.  This loader installs tiny stand-ins for
those modules (this directory) and a path hook that drops the ``default=`` keyword from
those three lines while the module source is being read (nothing is written or copied).
It exists so that tests/golden/make_golden.py can record golden vectors from the
reference's OWN code in the build container.  It is never used on the GPU box (the
reference is not present there) and never imported by the product package.
"""
import importlib.abc
import importlib.machinery
import os
import re
import sys

REFERENCE_ROOT = os.environ.get("TENSORSAT_REFERENCE", "/root/reference")
_DEFAULT_KW = re.compile(rb"(TypeVar\([^)\n]*?),\s*default=\w+\)")


class _Loader(importlib.machinery.SourceFileLoader):
    def get_data(self, path):
        data = super().get_data(path)
        if path.endswith(".py"):
            data = _DEFAULT_KW.sub(rb"\1)", data)
        return data

    def get_code(self, fullname):  # never use / write .pyc for patched sources
        path = self.get_filename(fullname)
        return self.source_to_code(self.get_data(path), path)


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path, target=None):
        if fullname != "tensorsat" and not fullname.startswith("tensorsat."):
            return None
        rel = fullname.replace(".", os.sep)
        base = os.path.join(REFERENCE_ROOT, rel)
        if os.path.isdir(base):
            init = os.path.join(base, "__init__.py")
            return importlib.machinery.ModuleSpec(
                fullname, _Loader(fullname, init), origin=init, is_package=True
            ) if os.path.exists(init) else None
        if os.path.exists(base + ".py"):
            return importlib.machinery.ModuleSpec(
                fullname, _Loader(fullname, base + ".py"), origin=base + ".py"
            )
        return None


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "tensorsat"))


def load():
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())
    import tensorsat  # noqa: F401

    for spec in (sys.modules["tensorsat"].__spec__,):
        if spec.submodule_search_locations is None:
            spec.submodule_search_locations = [os.path.join(REFERENCE_ROOT, "tensorsat")]
    return tensorsat
