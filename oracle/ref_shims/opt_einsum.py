"""Stand-in for ``opt_einsum`` (test infrastructure; see hashcons.py here).

``contract(a, wa, b, wb, wo)`` in interleaved form on two operands is numpy's own
einsum in the operands' dtype (the reference calls it at lang/fin_rel.py:276,283).
Labels are remapped to 0..51 because numpy's interleaved form caps labels at 52.
"""
import numpy as np


def contract(*args, **kw):
    *pairs, out = args
    ops, subs = pairs[0::2], pairs[1::2]
    labels = {}
    for s in list(subs) + [out]:
        for w in s:
            labels.setdefault(w, len(labels))
    flat = []
    for o, s in zip(ops, subs):
        flat += [o, [labels[w] for w in s]]
    return np.einsum(*flat, [labels[w] for w in out])


def contract_path(*a, **k):
    raise NotImplementedError("not needed for golden generation")


class paths:  # referenced only in type positions
    class PathOptimizer:
        pass
