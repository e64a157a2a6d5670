"""
Native front end: DIMACS text / clause lists -> flat relation network -> contraction tree -> device plan, without
building Python diagram objects (C-ABI: ``tsb_network_*`` / ``tsb_tree_*`` /
``tsb_plan_create_from_tree`` in include/tsb.h, implementation ``csrc/tsb_frontend.cu``).

For CNF inputs it replaces ``CNFInstance.from_dimacs`` / ``inputs`` / ``diagram`` (tensorsat/lib/sat.py:80-100,165-206,
248-262), ``>>`` + ``flatten`` (tensorsat/diagrams.py:387-461,1452-1481) and the tree search the reference delegates to
cotengra (tensorsat/contractions/cotengra.py:78-85): building the Python diagram of the n=200 benchmark instance takes
half a minute and the Python tree search seconds more; here the whole front end takes well under a second.  Wire and
slot numbering are exactly those of ``(cnf.inputs(x) >> cnf.diagram()).flatten()`` (pinned by tests against the
reference's wiring digests), so plans, trees and results are interchangeable with the mirror classes.

``is_satisfiable`` / ``count_models`` are the one-call forms of examples/test_sat_solver.py:5-9.
"""

from __future__ import annotations

import ctypes as C
from collections.abc import Sequence
from typing import Any

import numpy as np

from . import _native

LEAF_KINDS = ("bit_0", "bit_1", "bit_unk", "not_", "or_")


class _NetworkInfo(C.Structure):
    _fields_ = [("n_vars", C.c_int32), ("n_clauses", C.c_int32), ("n_leaves", C.c_int32), ("n_wires", C.c_int32),
                ("n_out", C.c_int32), ("reserved", C.c_int32), ("total_leaf_rank", C.c_int64), ("total_leaf_bytes", C.c_int64)]


class _TreeInfo(C.Structure):
    _fields_ = [("n_leaves", C.c_int32), ("n_steps", C.c_int32), ("n_sliced", C.c_int32), ("reserved", C.c_int32),
                ("log2_width", C.c_double), ("log2_width_unsliced", C.c_double), ("log2_ops", C.c_double), ("log2_write", C.c_double)]


def _cstr(s: str | None) -> Any:
    return None if s is None else C.c_char_p(s.encode())


class NativeTree:
    """Owner of a ``tsb_tree`` (SSA path + sliced wires + cost-model statistics)."""

    def __init__(self, handle: C.c_void_p) -> None:
        self._h = handle
        info = _TreeInfo()
        _native._check(_native.lib().tsb_tree_info(self._h, C.byref(info)))
        self.info = {f: getattr(info, f) for f, _ in _TreeInfo._fields_ if f != "reserved"}
        ssa = (C.c_int32 * max(1, 2 * info.n_steps))()
        sl = (C.c_int32 * max(1, info.n_sliced))()
        _native._check(_native.lib().tsb_tree_arrays(self._h, ssa, sl))
        self.ssa_path = tuple((int(ssa[2 * i]), int(ssa[2 * i + 1])) for i in range(info.n_steps))
        self.sliced_wires = tuple(int(sl[i]) for i in range(info.n_sliced))

    @property
    def num_slices(self) -> int:
        return 2 ** len(self.sliced_wires)

    def close(self) -> None:
        if self._h:
            _native.lib().tsb_tree_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class NativeNetwork:
    """Owner of a ``tsb_network``: the bintree relation network of a CNF, as flat wiring."""

    def __init__(self, handle: C.c_void_p) -> None:
        self._h = handle
        info = _NetworkInfo()
        _native._check(_native.lib().tsb_network_info(self._h, C.byref(info)))
        self.info = {f: getattr(info, f) for f, _ in _NetworkInfo._fields_ if f != "reserved"}
        self._staging: dict[str, Any] = {}

    @classmethod
    def from_dimacs(cls, text: str, inputs: str | None = None) -> "NativeNetwork":
        if not isinstance(text, str):
            raise TypeError("DIMACS code must be a string.")
        h = C.c_void_p()
        _native._check(_native.lib().tsb_network_from_dimacs(C.c_char_p(text.encode()), _cstr(inputs), C.byref(h)))
        return cls(h)

    @classmethod
    def from_clauses(cls, num_vars: int, clauses: Sequence[Sequence[int]], inputs: str | None = None) -> "NativeNetwork":
        sizes = [len(c) for c in clauses]
        lits = [int(l) for c in clauses for l in c]
        h = C.c_void_p()
        _native._check(_native.lib().tsb_network_from_clauses(int(num_vars), len(sizes), _native._i32(sizes), _native._i32(lits),
                                                              _cstr(inputs), C.byref(h)))
        return cls(h)

    def set_inputs(self, inputs: str) -> None:
        _native._check(_native.lib().tsb_network_set_inputs(self._h, _cstr(inputs)))

    def propagate(self, spec: str) -> tuple[bool, str]:
        """Unit propagation under the partial assignment ``spec``: ``(conflict, spec with implied values filled in)``.
        A conflict proves that no model extends ``spec``; no conflict proves nothing."""
        out = C.create_string_buffer(self.info["n_vars"] + 1)
        conflict = C.c_int32(0)
        _native._check(_native.lib().tsb_network_propagate(self._h, _cstr(spec), out, C.byref(conflict)))
        return bool(conflict.value), out.value.decode()

    # ---- the flat wiring, as the mirror classes would give it
    def arrays(self) -> tuple[tuple[tuple[int, ...], ...], tuple[int, ...], tuple[int, ...]]:
        """``(slot wire lists, leaf kinds, out wires)``."""
        n, tot = self.info["n_leaves"], self.info["total_leaf_rank"]
        ranks, wires, kinds = (C.c_int32 * max(1, n))(), (C.c_int32 * max(1, tot))(), (C.c_int32 * max(1, n))()
        outs = (C.c_int32 * max(1, self.info["n_out"]))()
        _native._check(_native.lib().tsb_network_arrays(self._h, ranks, wires, kinds, outs))
        inputs, q = [], 0
        for i in range(n):
            inputs.append(tuple(int(wires[q + j]) for j in range(ranks[i])))
            q += ranks[i]
        return tuple(inputs), tuple(int(kinds[i]) for i in range(n)), tuple(int(outs[i]) for i in range(self.info["n_out"]))

    def leaf_bytes(self) -> tuple[np.ndarray, np.ndarray]:
        """``(all leaf tensors concatenated as uint8, byte offset of every leaf)``."""
        buf = np.empty(max(1, self.info["total_leaf_bytes"]), dtype=np.uint8)
        offs = np.empty(max(1, self.info["n_leaves"]), dtype=np.int64)
        _native._check(_native.lib().tsb_network_leaf_tensors(self._h, buf.ctypes.data_as(C.POINTER(C.c_uint8)),
                                                              offs.ctypes.data_as(C.POINTER(C.c_int64))))
        return buf, offs

    def leaf_tensors(self) -> list[np.ndarray]:
        buf, offs = self.leaf_bytes()
        inputs, _, _ = self.arrays()
        return [buf[int(o): int(o) + (1 << len(ws))].reshape((2,) * len(ws)).copy() for o, ws in zip(offs, inputs)]

    # ---- tree and plan
    def search(self, *, trials: int = 2, seed: int = 0, target_log2_size: float = 30.0) -> NativeTree:
        h = C.c_void_p()
        _native._check(_native.lib().tsb_tree_search(self._h, int(trials), C.c_uint64(int(seed)), C.c_double(float(target_log2_size)),
                                                     C.byref(h)))
        return NativeTree(h)

    def plan(self, tree: NativeTree, mode: int = _native.MODE_SAT, *, batch_log2: int = -1, flags: int = 0) -> _native.Plan:
        h = C.c_void_p()
        _native._check(_native.lib().tsb_plan_create_from_tree(_native.ctx(), self._h, tree._h, int(mode), int(batch_log2),
                                                               C.c_uint32(int(flags)), C.byref(h)))
        return _native.Plan.from_handle(h, self.info["n_leaves"], mode)

    def contract(self, plan: _native.Plan, slice_range: tuple[int, int] | None = None) -> np.ndarray:
        """One H2D copy of the leaf tensors, one ``tsb_plan_execute``, the result back as a numpy array."""
        import torch

        buf, offs = self.leaf_bytes()
        key = str(_native.device())
        st = self._staging.get(key)
        if st is None:
            dev = torch.empty(buf.size, dtype=torch.uint8, device=_native.device())
            ptrs = (C.c_void_p * len(offs))(*[dev.data_ptr() + int(o) for o in offs])
            st = self._staging[key] = (dev, ptrs)
        dev, ptrs = st
        dev.copy_(torch.from_numpy(buf))
        out = _native.empty(plan.info["out_numel"], plan.mode)
        lo, hi = (0, plan.info["n_slices"]) if slice_range is None else slice_range
        plan.execute_ptrs(ptrs, out.data_ptr(), lo, hi)
        return out.cpu().numpy().reshape((2,) * self.info["n_out"])

    def close(self) -> None:
        if self._h:
            _native.lib().tsb_network_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def is_satisfiable(dimacs: str, *, trials: int = 2, seed: int = 0, target_log2_size: float = 30.0) -> bool:
    """SAT decision of a DIMACS instance on the device (examples/test_sat_solver.py:5-9), all inputs unknown."""
    net = NativeNetwork.from_dimacs(dimacs)
    net.close()
    net = NativeNetwork.from_dimacs(dimacs, "?" * net.info["n_vars"])
    tree = net.search(trials=trials, seed=seed, target_log2_size=target_log2_size)
    plan = net.plan(tree, _native.MODE_SAT)
    try:
        return bool(net.contract(plan).reshape(-1)[0])
    finally:
        plan.close()


def count_models(dimacs: str, *, trials: int = 2, seed: int = 0, target_log2_size: float = 30.0) -> int:
    """#SAT of a DIMACS instance on the device (exact int64; ``OverflowError`` if the bound exceeds int64)."""
    net = NativeNetwork.from_dimacs(dimacs)
    net.close()
    net = NativeNetwork.from_dimacs(dimacs, "?" * net.info["n_vars"])
    tree = net.search(trials=trials, seed=seed, target_log2_size=target_log2_size)
    plan = net.plan(tree, _native.MODE_COUNT)
    try:
        if plan.info["may_overflow"]:
            raise OverflowError("Counting bound of this instance exceeds int64.")
        return int(net.contract(plan).reshape(-1)[0])
    finally:
        plan.close()
