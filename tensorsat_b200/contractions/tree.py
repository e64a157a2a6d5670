"""
Whole-tree contraction on the device: the drop-in for ``CotengraContraction``
(tensorsat/contractions/cotengra.py:51-131).

The reference asks cotengra for a tree (``ctg.array_contract_tree``, cotengra.py:78-85) and lets
``tree.contract(diagram.boxes, prefer_einsum=True)`` drive one ``einsum`` per step through
autoray (cotengra.py:116-128).  Here the tree (found by ``tensorsat_b200.planner`` or supplied
by the caller, e.g. exported from cotengra as an SSA path plus sliced indices) is compiled once
into a ``tsb_plan`` — global wire order, static memory plan, lazy slicing — and every
``contract`` call is one ``tsb_plan_execute`` on the current CUDA stream.  The plan is reused
across calls with different leaf values (the assignment search of
notebooks/02-random-ksat.ipynb:707-722).

Slices can be restricted to a range (``slice_range``) so that ranks of a multi-GPU job each
contract their share; see ``tensorsat_b200.distributed``.
"""

from __future__ import annotations

from collections.abc import Sequence
from typing import Any

from .. import _native
from ..diagrams import Box, Diagram, TensorLikeBox, TensorLikeType, Wiring
from ..lang.fin_rel import MODE_COUNT, CudaCountRel, CudaFinRel, _CudaTensorBox
from ..planner import ContractionTree, Network
from ..planner.search import find_tree, find_tree_elimination
from .abc import Contraction

try:   # CPython helper of the binding (leaf packing for contract_from_host); optional, built by tensorsat_b200.build
    from .. import _tsb_host
except ImportError:   # pragma: no cover
    _tsb_host = None


class TreeContraction(Contraction):
    """A contraction that executes a whole contraction tree as one device plan."""

    def __init__(
        self,
        box_class: type,
        wiring: Wiring,
        *,
        tree: ContractionTree | None = None,
        ssa_path: Sequence[Sequence[int]] | None = None,
        sliced_wires: Sequence[int] = (),
        optimize: str = "auto",
        trials: int = 8,
        seed: int = 0,
        target_log2_size: float | None = None,
        wire_is_free: Sequence[bool] | None = None,
        batch_log2: int = -1,
        force_generic: bool = False,
        bit_packing: bool = True,
    ) -> None:
        super().__init__(box_class, wiring)
        if not issubclass(box_class, _CudaTensorBox):
            raise TypeError("TreeContraction executes on the device: box_class must be CudaFinRel or CudaCountRel.")
        if wiring.num_wires == 0:
            raise ValueError("Cannot define contraction for empty wiring.")
        if wiring.num_slots == 0:
            raise ValueError("Cannot define contraction for a wiring without slots.")
        for t in wiring.wire_types:
            if not isinstance(t, TensorLikeType):
                raise ValueError("Wiring must have tensor-like wire types.")
        net = Network.from_wiring(wiring)
        # repeated output wires: contract onto the distinct wires, then rewire (simple.py:287-289 does the same)
        self._distinct_out = tuple(dict.fromkeys(wiring.out_wires))
        if len(self._distinct_out) != len(wiring.out_wires):
            net = Network(net.inputs, self._distinct_out, net.size_dict)
        if tree is not None:
            if tree.network != net:
                raise ValueError("The tree was built for a different wiring.")
        elif ssa_path is not None:
            tree = ContractionTree(net, ssa_path, sliced_wires)
        else:
            if optimize == "greedy":
                tree = find_tree(net, trials=trials, seed=seed)
            else:
                a = find_tree_elimination(net, trials=trials, seed=seed)
                b = find_tree(net, trials=max(1, trials // 2), seed=seed)
                sa, sb = a.stats(), b.stats()
                tree = a if (sa.log2_max_size, sa.log2_ops) <= (sb.log2_max_size, sb.log2_ops) else b
            if sliced_wires:
                tree = tree.with_sliced(sliced_wires)
        if target_log2_size is not None and tree.width() > target_log2_size:
            tree = tree.slice_to(target_log2_size)
        self._tree = tree
        self._wire_is_free = None if wire_is_free is None else [bool(x) for x in wire_is_free]
        self._batch_log2 = int(batch_log2)
        self._flags = (_native.PLAN_FORCE_GENERIC if force_generic else 0) | (0 if bit_packing else _native.PLAN_NO_PACK)
        self._plan: _native.Plan | None = None

    @property
    def tree(self) -> ContractionTree:
        """Current contraction tree (cotengra.py:88-91)."""
        return self._tree

    @tree.setter
    def tree(self, tree: ContractionTree) -> None:
        """Replace the tree (e.g. after slicing it); the device plan is rebuilt on the next call."""
        if tree.network != self._tree.network:
            raise ValueError("The tree was built for a different wiring.")
        self._tree = tree
        self._plan = None

    @property
    def path(self) -> tuple[tuple[int, int], ...]:
        """Current contraction path, linear convention (cotengra.py:93-98)."""
        return self._tree.linear_path()

    @property
    def mode(self) -> int:
        return self.box_class._MODE

    @property
    def num_slices(self) -> int:
        return self._tree.num_slices

    def _accepts(self, cls: type) -> bool:
        # 0/1 relation leaves are valid inputs of either semiring
        return issubclass(cls, self._box_class) or issubclass(cls, CudaFinRel)

    def plan(self) -> _native.Plan:
        if self._plan is None:
            t = self._tree
            self._plan = _native.Plan(
                t.network.inputs, t.network.output, t.network.size_dict, t.ssa_path, t.sliced_wires,
                mode=self.mode, wire_is_free=self._wire_is_free, batch_log2=self._batch_log2, flags=self._flags,
            )
        return self._plan

    def _leaf_tensor(self, box: Box) -> Any:
        if isinstance(box, CudaFinRel):
            return box.device_tensor
        if isinstance(box, CudaCountRel):
            if not box.is_binary():
                # a plan stages its leaves as 0/1 bytes and sizes every intermediate from that bound
                raise ValueError("TreeContraction counts over 0/1 relation tensors: a leaf box has entries other "
                                 "than 0 and 1 (contract weighted tensors step by step with SimpleContraction).")
            t = box.device_tensor
            u8 = _native.empty(t.numel(), 0)
            _native.widen(t.reshape(-1), u8)
            return u8
        raise TypeError(f"Unsupported leaf box class {type(box)}.")

    def contract_raw(self, diagram: Diagram, slice_range: tuple[int, int] | None = None) -> Any:
        """Validates and contracts; returns the flat result tensor (u8 / int64) on the device."""
        self.validate(diagram)
        plan = self.plan()
        if plan.info["may_overflow"] and self.mode == MODE_COUNT:
            raise OverflowError("Counting bound of this network exceeds int64; mark determined wires "
                                "with wire_is_free or count a smaller instance.")
        leaves = [self._leaf_tensor(b) for b in diagram.boxes]
        out = _native.empty(plan.info["out_numel"], self.mode)
        lo, hi = (0, plan.info["n_slices"]) if slice_range is None else slice_range
        plan.execute(leaves, out, lo, hi)
        return out

    def contract_from_host(self, leaves: Sequence[Any], slice_range: tuple[int, int] | None = None, *,
                           group: Any = None, all_reduce: bool = False) -> Any:
        """End-to-end call on HOST buffers: ``leaves`` are the uint8 0/1 numpy arrays of the boxes (in
        slot order).  They are packed into one pinned staging buffer, copied to the device with one
        H2D transfer, contracted, and the result is read back (D2H) as a numpy array.  With ``all_reduce`` the
        partial result of this rank's ``slice_range`` is combined across the ranks of ``group`` on the device
        (one collective: max / sum) before the read-back."""
        import numpy as np
        import torch

        plan = self.plan()
        if len(leaves) != plan.n_leaves:
            raise ValueError(f"Expected {plan.n_leaves} leaf arrays, got {len(leaves)}.")
        # staging buffers live on one device: keyed by the device that is current for this call
        stagings = self.__dict__.setdefault("_stagings", {})
        dev_key = str(_native.device())
        st = stagings.get(dev_key)
        if st is None:
            sizes = [int(np.asarray(a).size) for a in leaves]
            total = int(sum(sizes))
            pin = torch.empty(total, dtype=torch.uint8)
            if _native._device == "cuda":
                pin = pin.pin_memory()
            dev = torch.empty(total, dtype=torch.uint8, device=_native.device())
            res = torch.empty(plan.info["out_numel"], dtype=torch.uint8 if self.mode == 0 else torch.int64)
            if _native._device == "cuda":
                res = res.pin_memory()
            offs = np.concatenate([[0], np.cumsum(sizes)[:-1]])
            import ctypes

            ptrs = (ctypes.c_void_p * len(leaves))(*[dev.data_ptr() + int(o) for o in offs])
            st = stagings[dev_key] = (pin, dev, res, sizes, ptrs)
        pin, dev, res, sizes, ptrs = st
        # one pass over the leaves: C-order bytes of every array, joined (uint8 / bool arrays are taken as they
        # are; anything else, or a leaf of another size, shows as a length mismatch and takes the checked path)
        pin_np = pin.numpy()
        if _tsb_host is not None and _tsb_host.pack_into(leaves, pin_np) == pin.numel():
            pass   # one pass through the buffer protocol, straight into the pinned buffer (csrc/tsb_pyhost.c)
        else:
            try:
                blob = bytearray().join(leaves)   # buffer protocol: no per-leaf bytes object (C-contiguous arrays only)
            except (TypeError, BufferError, ValueError):   # not buffers, or not contiguous
                blob = b""
            if len(blob) != pin.numel():
                if [int(np.asarray(a).size) for a in leaves] != sizes:
                    raise ValueError("Leaf sizes differ from those of the first call.")
                blob = np.concatenate([np.asarray(a, dtype=np.uint8).reshape(-1) for a in leaves]).tobytes()
            pin_np[:] = np.frombuffer(blob, dtype=np.uint8)
        lo, hi = (0, plan.info["n_slices"]) if slice_range is None else slice_range
        if not all_reduce:
            # one C-ABI call on host buffers: H2D of the packed leaves, the execution, D2H of the result (tsb_plan_execute_host)
            plan.execute_host(pin.data_ptr(), pin.numel(), res.data_ptr(), res.numel() * res.element_size(), lo, hi)
        else:
            dev.copy_(pin, non_blocking=True)
            out = _native.empty(plan.info["out_numel"], self.mode)
            plan.execute_ptrs(ptrs, out.data_ptr(), lo, hi)
            from ..distributed import all_reduce_result

            all_reduce_result(out, self.mode, group)
            res.copy_(out, non_blocking=False)
        # the plan contracts onto the DISTINCT output wires; repeated output wires tie axes to a diagonal
        # (what contract() does with _rewire; simple.py:287-289)
        dims = tuple(self.wiring.wire_dims[w] for w in self._distinct_out)
        self.h2d_bytes, self.d2h_bytes = int(pin.numel()), int(res.numel() * res.element_size())
        arr = res.numpy().reshape(dims).copy()
        if self._distinct_out != tuple(self.wiring.out_wires):
            ports = [self._distinct_out.index(w) for w in self.wiring.out_wires]
            full = np.zeros(tuple(dims[p] for p in ports), dtype=arr.dtype)
            for idx in np.ndindex(*dims):
                full[tuple(idx[p] for p in ports)] = arr[idx]
            arr = full
        return arr

    def contract(self, diagram: Diagram, *, slice_range: tuple[int, int] | None = None, **_ignored: Any) -> Box:
        """Validates and contracts the given diagram (cotengra.py:100-128).

        Keyword arguments of the reference (``order``, ``autojit``, ``progbar``) are accepted and ignored.
        """
        out = self.contract_raw(diagram, slice_range)
        dims = tuple(self.wiring.wire_dims[w] for w in self._distinct_out)
        box = self.box_class._new(dims, dev=out)
        if self._distinct_out != tuple(self.wiring.out_wires):
            box = box._rewire([self._distinct_out.index(w) for w in self.wiring.out_wires])
        return box

    def __repr__(self) -> str:
        return "<TreeContraction>"
