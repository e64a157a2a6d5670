"""Stand-in for the un-installed third-party ``hashcons`` package (test infrastructure).

Only what the reference touches (tensorsat/diagrams.py:156-180, lang/fin_rel.py:82-93):
``InstanceStore().instance(cls, key)`` is a context manager yielding the interned
instance or ``None``; ``register(obj)`` interns the object under the pending key.
Used ONLY by tests/golden/make_golden.py, in the build container, to import the real
reference from /root/reference and record golden vectors.  Never imported by the product.
"""
from contextlib import contextmanager


class InstanceStore:
    def __init__(self):
        self._d = {}
        self._pending = []

    @contextmanager
    def instance(self, cls, key):
        k = (cls, key)
        self._pending.append(k)
        try:
            yield self._d.get(k)
        finally:
            self._pending.pop()

    def register(self, obj):
        self._d[self._pending[-1]] = obj
