"""Stand-in for ``autoray`` (test infrastructure; see hashcons.py here).

Records backend registrations (tensorsat/diagrams.py:684-686) and supplies the
``shape.register`` decorator used at tensorsat/_autoray.py:69 so that the reference's
own ``einsum/transpose/stack`` shim functions get defined and can be driven directly.
"""
_BACKENDS = {}


def register_backend(cls, name):
    _BACKENDS[cls] = name


class _Shape:
    def __init__(self):
        self.registry = {}

    def register(self, name):
        def deco(fn):
            self.registry[name] = fn
            return fn
        return deco


shape = _Shape()
