"""
Explicit-path contraction, one ``Box._contract2`` call per step
(tensorsat/contractions/simple.py:75-295).  Every step is one ``tsb_contract2`` launch on
device-resident boxes; nothing is planned ahead, so this is the step-wise route through the
plug-in boundary (what cotengra does through autoray).  For whole-tree execution with a
memory plan and lazy slicing use :class:`~tensorsat_b200.contractions.tree.TreeContraction`.
"""

from __future__ import annotations

from collections.abc import Sequence

from ..diagrams import Box, Diagram, TensorLikeBox, TensorLikeType, Wiring
from .abc import Contraction

ContractionPath = tuple[tuple[int, int], ...]
"""Linear path convention of ``opt_einsum.contract_path`` (simple.py:38-46): the indices of a
pair select boxes in the current list; both are removed and the result is appended."""


class SimpleContraction(Contraction):
    """A simple contraction based on an explicit contraction path."""

    def __init__(self, box_class: type, wiring: Wiring, path: Sequence[tuple[int, int]]) -> None:
        super().__init__(box_class, wiring)
        if not issubclass(box_class, TensorLikeBox):
            raise TypeError("box_class must be a tensor-like box class.")
        path = tuple((int(a), int(b)) for a, b in path)
        if wiring.num_wires == 0:
            raise ValueError("Cannot define contraction for empty wiring.")
        for t in wiring.wire_types:
            if not isinstance(t, TensorLikeType):
                raise ValueError("Wiring must have tensor-like wire types.")
        n = wiring.num_slots
        if not path and n >= 2:
            raise ValueError("Contraction path can only be empty for number of wiring slots <= 1.")
        if len(path) >= n and n > 0:
            raise ValueError(
                "Path is too long for the given wiring:"
                f" there are {len(path)} contractions, but {n} slots to contract."
            )
        if len(path) != max(n - 1, 0):
            raise ValueError(
                "Path does not fully contract the given wiring:"
                f" there are {len(path)} contractions, but {n} slots to contract."
            )
        for idx, (lhs, rhs) in enumerate(path):
            if not 0 <= lhs < n - idx:
                raise ValueError(f"Invalid lhs for contraction {(lhs, rhs) = } at {idx = }.")
            if not 0 <= rhs < n - idx or rhs == lhs:
                raise ValueError(f"Invalid rhs for contraction {(lhs, rhs) = } at {idx = }.")
        self._path = path
        # arguments of the contract2 calls (simple.py:134-194)
        args: list[tuple[int, tuple[int, ...], int, tuple[int, ...], list[int]]] = []
        if path:
            wires_list = list(wiring.slot_wires_list)
            for lhs_idx, rhs_idx in path:
                lhs_wires = wires_list.pop(lhs_idx)
                rhs_wires = wires_list.pop(rhs_idx - int(lhs_idx < rhs_idx))
                common = set(lhs_wires) & set(rhs_wires)
                if common:
                    res = (
                        tuple(w for w in lhs_wires if w not in common)
                        + tuple(w for w in lhs_wires if w in common)
                        + tuple(w for w in rhs_wires if w not in common)
                    )
                else:
                    res = lhs_wires + rhs_wires
                wires_list.append(res)
                args.append((lhs_idx, lhs_wires, rhs_idx, rhs_wires, list(res)))
            final = args[-1][-1]
            # drop each non-output wire from the result of the LAST call in which it features
            for w in sorted(set(final) - set(wiring.out_wires)):
                for *_, res_wires in reversed(args):
                    if w in res_wires:
                        res_wires.remove(w)
                        break
            self._box_out_wires: tuple[int, ...] = tuple(final)
        elif wiring.num_slots == 0:
            self._box_out_wires = ()
        else:
            self._box_out_wires = wiring.slot_wires_list[0]
        self._contract2_args = tuple((a, lw, b, rw, tuple(res)) for a, lw, b, rw, res in args)
        self._dangling_wires = tuple(sorted(wiring.dangling_wires))

    @property
    def path(self) -> ContractionPath:
        return self._path

    @property
    def contract2_args(self):
        return self._contract2_args

    def contract(self, diagram: Diagram) -> Box:
        """Validates and contracts the given diagram (simple.py:251-292)."""
        self.validate(diagram)
        box_class, wiring = self.box_class, self.wiring
        boxes = list(diagram.boxes)
        box = None
        if boxes:
            for lhs_idx, lhs_wires, rhs_idx, rhs_wires, res_wires in self._contract2_args:
                lhs = boxes.pop(lhs_idx)
                rhs = boxes.pop(rhs_idx - int(lhs_idx < rhs_idx))
                boxes.append(box_class._contract2(lhs, lhs_wires, rhs, rhs_wires, res_wires))
            assert len(boxes) == 1
            box = boxes[0]
        box_out_wires = self._box_out_wires
        if self._dangling_wires:
            box_out_wires = box_out_wires + self._dangling_wires
            spiders = [box_class._spider(wiring.wire_types[w], 1) for w in self._dangling_wires]
            box = box_class.prod(([box] if box is not None else []) + spiders)
        if box is None:
            raise ValueError("Nothing to contract.")
        if wiring.out_wires != box_out_wires:
            box = box._rewire([box_out_wires.index(w) for w in wiring.out_wires])
        return box

    def __repr__(self) -> str:
        return "<SimpleContraction>"
