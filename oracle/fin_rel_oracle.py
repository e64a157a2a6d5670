"""
ORACLE (test infrastructure, not product code): CPU restatement in numpy of the reference's
contraction arithmetic.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module, and only as the checker /
the CPU baseline — never as a fallback of the product path.

Parity status: PINNED.  Every function below is checked (tests/test_oracle_golden.py)
against golden vectors recorded from the reference's OWN code, imported in the build
container by ``oracle/ref_shims/load_reference.py`` and driven by
``tests/golden/make_golden.py`` (fixtures under ``tests/golden/``), and against the
known answers the reference's notebooks/tests hold.

Each function cites the reference lines it follows (paths relative to the reference repo).
"""

from __future__ import annotations

from collections import Counter
from itertools import product
from math import prod
from typing import Iterable, Sequence

import numpy as np

_NP_EINSUM_MAX_LABELS = 52


def to_safe_contraction_dtype(contraction_size: int, *tensors: np.ndarray) -> tuple[np.ndarray, ...]:
    """tensorsat/lang/fin_rel.py:145-159 — widen so that the integer sum cannot wrap."""
    assert contraction_size >= 0
    if contraction_size < 256:
        return tensors
    if contraction_size < 65536:
        dt = np.dtype("uint16")
    elif contraction_size < 2**32:
        dt = np.dtype("uint32")
    else:
        dt = np.dtype("uint64")
    return tuple(t.astype(dt) for t in tensors)


def einsum2(
    lhs: np.ndarray,
    lhs_wires: Sequence[int],
    rhs: np.ndarray,
    rhs_wires: Sequence[int],
    out_wires: Sequence[int],
) -> np.ndarray:
    """``opt_einsum.contract(lhs, lw, rhs, rw, ow)`` on two operands (fin_rel.py:276,283).

    For two operands that is numpy's einsum in the operands' dtype.  numpy's interleaved
    form handles at most 52 distinct labels, so labels are renumbered densely first; beyond
    52 the same sum is restated via transpose + reshape + matmul (integer matmul, same dtype).
    """
    labels: dict[int, int] = {}
    for w in list(lhs_wires) + list(rhs_wires) + list(out_wires):
        labels.setdefault(w, len(labels))
    if len(labels) <= _NP_EINSUM_MAX_LABELS:
        return np.einsum(
            lhs,
            [labels[w] for w in lhs_wires],
            rhs,
            [labels[w] for w in rhs_wires],
            [labels[w] for w in out_wires],
        )
    return _einsum2_matmul(lhs, lhs_wires, rhs, rhs_wires, out_wires)


def _einsum2_matmul(lhs, lhs_wires, rhs, rhs_wires, out_wires) -> np.ndarray:
    lw, rw, ow = list(lhs_wires), list(rhs_wires), list(out_wires)
    if len(set(lw)) != len(lw) or len(set(rw)) != len(rw) or len(set(ow)) != len(ow):
        raise NotImplementedError("repeated labels with more than 52 distinct labels")
    dim = {w: d for w, d in list(zip(lw, lhs.shape)) + list(zip(rw, rhs.shape))}
    so = set(ow)
    # lone wires that are summed on one side only
    lhs = lhs.sum(axis=tuple(i for i, w in enumerate(lw) if w not in so and w not in rw), dtype=lhs.dtype)
    lw = [w for w in lw if w in so or w in rw]
    rhs = rhs.sum(axis=tuple(i for i, w in enumerate(rw) if w not in so and w not in lw), dtype=rhs.dtype)
    rw = [w for w in rw if w in so or w in lw]
    B = [w for w in lw if w in rw and w in so]
    K = [w for w in lw if w in rw and w not in so]
    M = [w for w in lw if w not in rw]
    N = [w for w in rw if w not in lw]
    a = np.transpose(lhs, [lw.index(w) for w in B + M + K]).reshape(
        prod(dim[w] for w in B), prod(dim[w] for w in M), prod(dim[w] for w in K)
    )
    b = np.transpose(rhs, [rw.index(w) for w in B + K + N]).reshape(
        prod(dim[w] for w in B), prod(dim[w] for w in K), prod(dim[w] for w in N)
    )
    c = np.matmul(a, b).reshape([dim[w] for w in B + M + N])
    cur = B + M + N
    return np.transpose(c, [cur.index(w) for w in ow])


def rewire_array_exact(a: np.ndarray, out_ports: Sequence[int]) -> np.ndarray:
    """What ``rewire_array`` is meant to compute: ``b[i_out] = sum over dropped ports of a``,
    duplicated ports tied to the diagonal, surviving ports placed where ``out_ports`` says."""
    out_ports = list(out_ports)
    distinct = list(dict.fromkeys(out_ports))
    red = np.einsum(a, list(range(a.ndim)), distinct) if a.ndim else a
    b = np.zeros(tuple(a.shape[p] for p in out_ports), dtype=a.dtype)
    for idx in np.ndindex(*[a.shape[p] for p in distinct]):
        pos = dict(zip(distinct, idx))
        b[tuple(pos[p] for p in out_ports)] = red[idx]
    return b


def rewire_array(a: np.ndarray, out_ports: Sequence[int]) -> np.ndarray:
    """tensorsat/_utils/misc.py:104-123 verbatim — transpose with port duplication and discarding.

    Reference quirk R2: the slice ``a[a_idxs]`` keeps the surviving ports in ``a``'s order while
    ``b[b_idxs]`` expects them in ``out_ports`` order, so the function is only right when the
    surviving (non-repeated, non-dropped) ports keep their relative order; otherwise it adds
    un-transposed data (all dims equal) or raises ``ValueError`` (dims differ).  Recorded in
    tests/golden/ops.json; use :func:`rewire_array_exact` for the intended semantics.
    """
    a_shape = a.shape
    a_ports = range(len(a_shape))
    b_shape = tuple(a_shape[port] for port in out_ports)
    b = np.zeros(b_shape, dtype=a.dtype)
    repeat_ports = sorted(p for p, c in Counter(out_ports).items() if c > 1)
    repeat_ports += sorted(set(a_ports) - set(out_ports))
    port_i = {p: i for i, p in enumerate(repeat_ports)}
    for _idxs in product(*(range(a_shape[p]) for p in repeat_ports)):
        a_idxs = tuple(_idxs[port_i[p]] if p in port_i else slice(None) for p in a_ports)
        b_idxs = tuple(_idxs[port_i[p]] if p in port_i else slice(None) for p in out_ports)
        b[b_idxs] += a[a_idxs]
    return b


def clamp(x: np.ndarray, *, ref_quirk_mod256: bool = False) -> np.ndarray:
    """The boolean-semiring clamp.

    ``ref_quirk_mod256=True`` is the reference verbatim (quirk R1): ``np.sign(x, dtype=np.uint8)``
    (fin_rel.py:279,287,392), which casts the widened sum to uint8 *before* the sign, so
    a non-zero multiple of 256 clamps to 0.  The default is the mathematical ``x > 0``.
    The two agree whenever no summed count is a non-zero multiple of 256.
    """
    if ref_quirk_mod256:
        return np.sign(x, dtype=np.uint8, casting="unsafe")
    return (x != 0).astype(np.uint8)


def contract2(
    lhs: np.ndarray,
    lhs_wires: Sequence[int],
    rhs: np.ndarray,
    rhs_wires: Sequence[int],
    out_wires: Sequence[int],
    *,
    mode: str = "sat",
    ref_quirk_mod256: bool = False,
    ref_quirk_rewire: bool = False,
) -> np.ndarray:
    """``FinRel._contract2`` (fin_rel.py:255-288).  ``mode="count"``: same sum in int64, no clamp
    (the #SAT prototype, notebooks/prototypes/2025-02-13-#sat-tensor-contraction.ipynb:215,244).

    With both ``ref_quirk_*`` flags set this is the reference bit for bit; the defaults give
    the semantics the reference intends (see :func:`clamp`, :func:`rewire_array`)."""
    lhs_wires, rhs_wires, out_wires = list(lhs_wires), list(rhs_wires), list(out_wires)
    if mode == "count":
        lhs_t, rhs_t = lhs.astype(np.int64), rhs.astype(np.int64)
    else:
        contraction_size = prod(
            dim
            for dim, w in zip(lhs.shape, lhs_wires)
            if w in (set(lhs_wires) & set(rhs_wires)) - set(out_wires)
        )
        lhs_t, rhs_t = to_safe_contraction_dtype(contraction_size, lhs, rhs)
    out_wireset = set(out_wires)
    if len(out_wireset) == len(out_wires):
        res = einsum2(lhs_t, lhs_wires, rhs_t, rhs_wires, out_wires)
    else:
        einsum_out_wires = sorted(out_wireset)
        rewire_out_ports = [einsum_out_wires.index(w) for w in out_wires]
        res = einsum2(lhs_t, lhs_wires, rhs_t, rhs_wires, einsum_out_wires)
        res = (rewire_array if ref_quirk_rewire else rewire_array_exact)(res, rewire_out_ports)
    if mode == "count":
        return np.array(res, order="C")
    return clamp(res, ref_quirk_mod256=ref_quirk_mod256)


def rewire(
    tensor: np.ndarray,
    out_ports: Sequence[int],
    *,
    mode: str = "sat",
    ref_quirk_mod256: bool = False,
    ref_quirk_rewire: bool = False,
) -> np.ndarray:
    """``FinRel._rewire`` (fin_rel.py:379-393)."""
    _rewire_array = rewire_array if ref_quirk_rewire else rewire_array_exact
    out_ports = list(out_ports)
    out_portset = frozenset(out_ports)
    if len(out_portset) == len(out_ports) == tensor.ndim:
        return np.transpose(tensor, out_ports)
    if mode == "count":
        return _rewire_array(tensor.astype(np.int64), out_ports)
    discarded = frozenset(range(tensor.ndim)) - out_portset
    if discarded:
        contraction_size = prod(d for p, d in enumerate(tensor.shape) if p in discarded)
        (tensor,) = to_safe_contraction_dtype(contraction_size, tensor)
    res = _rewire_array(tensor, out_ports)
    return clamp(res, ref_quirk_mod256=ref_quirk_mod256)


def spider(size: int, num_ports: int, dtype=np.uint8) -> np.ndarray:
    """``FinRel._spider`` (fin_rel.py:290-298)."""
    tensor = np.zeros(shape=(size,) * num_ports, dtype=dtype)
    for i in range(size):
        tensor[(i,) * num_ports] = 1
    return tensor


def scalar(value: bool, dtype=np.uint8) -> np.ndarray:
    """``FinRel._scalar`` (fin_rel.py:300-302)."""
    return np.array(int(value), dtype=dtype)


def stack(tensors: Sequence[np.ndarray], port: int = 0) -> np.ndarray:
    """``FinRel._stack`` (fin_rel.py:304-308)."""
    return np.stack(tuple(tensors), axis=port)


def getitem(tensor: np.ndarray, idxs: tuple) -> np.ndarray:
    """``FinRel._getitem`` (fin_rel.py:342-343)."""
    return tensor[idxs]


def union(a: np.ndarray, b: np.ndarray, *, mode: str = "sat") -> np.ndarray:
    """``FinRel._union`` (fin_rel.py:345-346); counting mode adds (disjoint slices)."""
    return (a + b) if mode == "count" else (a | b)


def as_scalar(tensor: np.ndarray) -> bool:
    """``FinRel._as_scalar`` (fin_rel.py:472-473)."""
    return bool(np.all(tensor == 1))


def parse_einsum(contraction: str) -> tuple[list[int], list[int], list[int]]:
    """Letters -> wire ids exactly as tensorsat/_autoray.py:53-60 (first position in lhs+rhs)."""
    inputs, out = contraction.split("->")
    lhs_s, rhs_s = inputs.split(",")
    char_idxs = {c: i for i, c in reversed(list(enumerate(lhs_s + rhs_s)))}
    return [char_idxs[c] for c in lhs_s], [char_idxs[c] for c in rhs_s], [char_idxs[c] for c in out]


# --------------------------------------------------------------------------------------
# Whole-network execution: the stand-in for ``tree.contract(diagram.boxes, prefer_einsum=True)``
# (tensorsat/contractions/cotengra.py:120-126): one einsum per tree step through the
# reference arithmetic above, slices gathered by union (fin_rel.py:345) / stack (:304).
# --------------------------------------------------------------------------------------
def execute_tree(
    inputs: Sequence[Sequence[int]],
    output: Sequence[int],
    leaves: Sequence[np.ndarray],
    ssa_path: Iterable[tuple[int, int]],
    sliced_wires: Sequence[int] = (),
    *,
    mode: str = "sat",
    ref_quirk_mod256: bool = False,
    slice_ids: Iterable[int] | None = None,
    size_dict: Sequence[int] | None = None,
) -> np.ndarray:
    """Contract a flat network along an SSA path; returns the tensor over ``output`` wires.

    With sliced wires, slice ``i`` fixes the sliced wires to the mixed-radix digits of ``i``
    (first sliced wire most significant); ``slice_ids`` restricts execution to a subset of
    slices (used for sampled-slice baselines); sliced output wires are not supported here.
    """
    inputs = [tuple(ws) for ws in inputs]
    output = tuple(output)
    sliced = list(sliced_wires)
    if set(sliced) & set(output):
        raise NotImplementedError("sliced output wires")
    ssa_path = list(ssa_path)
    L = len(inputs)
    dim: dict[int, int] = {}
    for ws, t in zip(inputs, leaves):
        for w, d in zip(ws, t.shape):
            dim[w] = d
    if size_dict is not None:
        for w in sliced:
            dim.setdefault(w, size_dict[w])
    nslices = prod(dim[w] for w in sliced) if sliced else 1
    ids = range(nslices) if slice_ids is None else slice_ids
    total = None
    for sid in ids:
        digits = {}
        rem = sid
        for w in reversed(sliced):
            digits[w] = rem % dim[w]
            rem //= dim[w]
        tensors: dict[int, np.ndarray] = {}
        wires: dict[int, tuple[int, ...]] = {}
        for i, (ws, t) in enumerate(zip(inputs, leaves)):
            idx = tuple(digits[w] if w in digits else slice(None) for w in ws)
            tensors[i] = getitem(np.asarray(t), idx)
            wires[i] = tuple(w for w in ws if w not in digits)
        count: Counter = Counter()
        for ws in wires.values():
            count.update(set(ws))
        count.update(set(output))
        if L == 1:
            (ws,) = wires.values()
            res = rewire(tensors[0], [ws.index(w) for w in output], mode=mode,
                         ref_quirk_mod256=ref_quirk_mod256)
        else:
            for s, (a, b) in enumerate(ssa_path):
                la, lb = wires.pop(a), wires.pop(b)
                ta, tb = tensors.pop(a), tensors.pop(b)
                sa, sb = set(la), set(lb)
                out_w = []
                for w in list(dict.fromkeys(la + lb)):
                    c = count[w] - (1 if (w in sa and w in sb) else 0)
                    if c >= 2:
                        out_w.append(w)
                        count[w] = c
                    else:
                        count[w] = 0
                if s == len(ssa_path) - 1:
                    out_w = list(output)
                tensors[L + s] = contract2(ta, la, tb, lb, out_w, mode=mode,
                                           ref_quirk_mod256=ref_quirk_mod256)
                wires[L + s] = tuple(out_w)
            (res,) = tensors.values()
        res = np.asarray(res)
        total = res if total is None else union(total, res, mode=mode)
    assert total is not None
    return total


# --------------------------------------------------------------------------------------
# Independent checker: brute-force enumeration of a CNF (no tensors involved).
# --------------------------------------------------------------------------------------
def cnf_brute_force(num_vars: int, clauses: Sequence[Sequence[int]]) -> np.ndarray:
    """Boolean vector over all 2**num_vars assignments (variable 1 = most significant bit):
    1 where every clause is satisfied.  Vectorised; practical for num_vars <= 24."""
    if num_vars > 26:
        raise ValueError("brute force limited to 26 variables")
    n = num_vars
    idx = np.arange(2**n, dtype=np.uint32)
    ok = np.ones(2**n, dtype=bool)
    for clause in clauses:
        sat = np.zeros(2**n, dtype=bool)
        for lit in clause:
            bit = (idx >> np.uint32(n - abs(lit))) & np.uint32(1)
            sat |= (bit == 1) if lit > 0 else (bit == 0)
        ok &= sat
    return ok.astype(np.uint8)
