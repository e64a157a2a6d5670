"""
Contraction trees over a flat :class:`Network`, their cost statistics and slicing.

This is the stand-in for what the reference obtains from cotengra
(``ctg.array_contract_tree`` at tensorsat/contractions/cotengra.py:78-85 and the
in-place ``slice_and_reconfigure`` of notebooks/02-random-ksat.ipynb:599-627): it is
host-side search, not part of the timed contraction phase.  The five statistics are the
ones the reference prints (notebooks/02-random-ksat.ipynb:284-290): number of
contractions, total ops, total write, max tensor size and peak live size.

Paths are kept in SSA form (leaf ids ``0..L-1``; step ``s`` creates id ``L+s``) and can
be converted to / from the reference's linear ``ContractionPath`` convention
(tensorsat/contractions/simple.py:38-46: pop both operands, append the result).
"""

from __future__ import annotations

import json
import math
from dataclasses import dataclass, field
from typing import Any, Iterable, Sequence

from .network import Network

SSAPath = tuple[tuple[int, int], ...]


def ssa_to_linear(path: Sequence[tuple[int, int]], num_leaves: int) -> tuple[tuple[int, int], ...]:
    """SSA ids -> positions in the shrinking list (tensorsat/contractions/simple.py:38-46)."""
    ids = list(range(num_leaves))
    out = []
    for s, (a, b) in enumerate(path):
        ia, ib = ids.index(a), ids.index(b)
        out.append((ia, ib))
        # reference semantics: pop lhs, then pop rhs - (lhs < rhs), append the result
        ids.pop(ia)
        ids.pop(ib - int(ia < ib))
        ids.append(num_leaves + s)
    return tuple(out)


def linear_to_ssa(path: Sequence[tuple[int, int]], num_leaves: int) -> SSAPath:
    ids = list(range(num_leaves))
    out = []
    for s, (ia, ib) in enumerate(path):
        if ia == ib:
            raise ValueError(f"Invalid contraction pair {(ia, ib)} at step {s}.")
        a = ids.pop(ia)
        b = ids.pop(ib - int(ia < ib))
        out.append((a, b))
        ids.append(num_leaves + s)
    return tuple(out)


@dataclass
class Step:
    lhs: int
    rhs: int
    out: int
    lhs_wires: frozenset[int]
    rhs_wires: frozenset[int]
    out_wires: frozenset[int]

    @property
    def kept_batch(self) -> frozenset[int]:
        return self.lhs_wires & self.rhs_wires & self.out_wires

    @property
    def summed(self) -> frozenset[int]:
        return (self.lhs_wires | self.rhs_wires) - self.out_wires


@dataclass
class TreeStats:
    num_contractions: int
    log2_ops: float  # log2 of sum over steps of 2 * prod(dims of all wires in the step)
    log2_write: float
    log2_max_size: float
    log2_peak_size: float
    num_slices: int
    sliced_wires: tuple[int, ...]

    def to_json(self) -> dict[str, Any]:
        return self.__dict__ | {"sliced_wires": list(self.sliced_wires)}


class ContractionTree:
    """A full pairwise contraction order for a network, with optional sliced wires."""

    def __init__(
        self,
        network: Network,
        ssa_path: Iterable[tuple[int, int]],
        sliced_wires: Iterable[int] = (),
    ) -> None:
        self.network = network
        self.ssa_path: SSAPath = tuple((int(a), int(b)) for a, b in ssa_path)
        self.sliced_wires: tuple[int, ...] = tuple(int(w) for w in sliced_wires)
        L = network.num_leaves
        if L == 0:
            raise ValueError("Cannot contract an empty network.")
        if len(self.ssa_path) != L - 1:
            raise ValueError(
                f"Path does not fully contract the network: {len(self.ssa_path)} "
                f"contractions for {L} tensors."
            )
        for w in self.sliced_wires:
            if not 0 <= w < network.num_wires:
                raise ValueError(f"Invalid sliced wire {w}.")
        if len(set(self.sliced_wires)) != len(self.sliced_wires):
            raise ValueError("Sliced wires must be distinct.")
        self.steps: list[Step] = self._build_steps()

    # -- structure ---------------------------------------------------------------
    def _build_steps(self) -> list[Step]:
        net = self.network
        L = net.num_leaves
        sliced = set(self.sliced_wires)
        sets: dict[int, frozenset[int]] = {
            i: frozenset(w for w in ws if w not in sliced) for i, ws in enumerate(net.inputs)
        }
        count: dict[int, int] = {}
        for s in sets.values():
            for w in s:
                count[w] = count.get(w, 0) + 1
        for w in set(net.output):
            if w not in sliced:
                count[w] = count.get(w, 0) + 1
        steps: list[Step] = []
        alive = set(range(L))
        for s, (a, b) in enumerate(self.ssa_path):
            if a == b or a not in alive or b not in alive:
                raise ValueError(f"Invalid contraction pair {(a, b)} at step {s}.")
            la, lb = sets[a], sets[b]
            out = set()
            for w in la | lb:
                c = count[w] - (1 if (w in la and w in lb) else 0)
                if c >= 2:
                    out.add(w)
                    count[w] = c
                else:
                    count[w] = 0
            alive.discard(a)
            alive.discard(b)
            alive.add(L + s)
            sets[L + s] = frozenset(out)
            steps.append(Step(a, b, L + s, la, lb, frozenset(out)))
        return steps

    @property
    def num_slices(self) -> int:
        return math.prod(self.network.size_dict[w] for w in self.sliced_wires)

    def linear_path(self) -> tuple[tuple[int, int], ...]:
        return ssa_to_linear(self.ssa_path, self.network.num_leaves)

    # -- statistics --------------------------------------------------------------
    def stats(self) -> TreeStats:
        dims = self.network.size_dict
        lg = [math.log2(d) for d in dims]
        ops = 0.0
        write = 0.0
        max_size = 0.0
        for ws in self.network.inputs:
            max_size = max(max_size, sum(lg[w] for w in ws if w not in self.sliced_wires))
        sizes: dict[int, float] = {}
        sliced = set(self.sliced_wires)
        for i, ws in enumerate(self.network.inputs):
            sizes[i] = 2.0 ** sum(lg[w] for w in set(ws) - sliced)
        live = sum(sizes.values())
        peak = live
        for st in self.steps:
            allw = st.lhs_wires | st.rhs_wires
            ops += 2.0 ** (1 + sum(lg[w] for w in allw))
            osz = sum(lg[w] for w in st.out_wires)
            write += 2.0**osz
            max_size = max(max_size, osz)
            sizes[st.out] = 2.0**osz
            live += sizes[st.out]
            peak = max(peak, live)
            live -= sizes.pop(st.lhs) + sizes.pop(st.rhs)
        n = self.num_slices
        return TreeStats(
            num_contractions=len(self.steps),
            log2_ops=math.log2(max(ops, 1.0) * n),
            log2_write=math.log2(max(write, 1.0) * n),
            log2_max_size=max_size,
            log2_peak_size=math.log2(max(peak, 1.0)),
            num_slices=n,
            sliced_wires=self.sliced_wires,
        )

    def width(self) -> float:
        return self.stats().log2_max_size

    # -- slicing -----------------------------------------------------------------
    def with_sliced(self, sliced_wires: Iterable[int]) -> "ContractionTree":
        return ContractionTree(self.network, self.ssa_path, sliced_wires)

    def slice_to(self, target_log2_size: float, *, max_slices_log2: int = 40) -> "ContractionTree":
        """Greedy slicing: add sliced wires until the largest intermediate fits.

        At each round the candidates are the wires of the current largest intermediates
        (never output wires); the wire whose removal gives the smallest total op count is
        taken (cotengra's SliceFinder does the same with randomisation).
        """
        out = set(self.network.output)
        tree = self
        while tree.width() > target_log2_size + 1e-9:
            if len(tree.sliced_wires) >= max_slices_log2:
                raise ValueError("Slicing budget exhausted before reaching the target size.")
            w_max = tree.width()
            lg = [math.log2(d) for d in self.network.size_dict]
            cands: set[int] = set()
            for st in tree.steps:
                if sum(lg[w] for w in st.out_wires) >= w_max - 1e-9:   # log2 size, any extents
                    cands |= st.out_wires
            done = set(tree.sliced_wires)
            for ws in self.network.inputs:   # a leaf can be the largest tensor left
                rest = set(ws) - done
                if sum(lg[w] for w in rest) >= w_max - 1e-9:
                    cands |= rest
            cands -= out
            if not cands:
                raise ValueError("No sliceable wire left in the largest intermediates.")
            best = None
            for w in sorted(cands):
                t = tree.with_sliced(tree.sliced_wires + (w,))
                st = t.stats()
                key = (st.log2_max_size, st.log2_ops)
                if best is None or key < best[0]:
                    best = (key, t)
            assert best is not None
            tree = best[1]
        return tree

    def slice_to_fast(self, target_log2_size: float, *, max_sliced: int = 4096) -> "ContractionTree":
        """Slicing for very wide trees (hundreds of sliced wires): each round slices the non-output wire
        that carries the most weight, where a wire's weight is the summed size of the intermediates above
        the target that hold it.  One pass over the steps per round instead of one tree rebuild per candidate."""
        out = set(self.network.output)
        lg = [math.log2(d) for d in self.network.size_dict]
        sliced: list[int] = list(self.sliced_wires)
        sets = [set(st.out_wires) for st in self.steps]
        while True:
            weight: dict[int, float] = {}
            worst = 0.0
            for ws in sets:
                sz = sum(lg[w] for w in ws)
                worst = max(worst, sz)
                if sz > target_log2_size + 1e-9:
                    bonus = 2.0 ** min(sz - target_log2_size, 60.0)
                    for w in ws:
                        if w not in out:
                            weight[w] = weight.get(w, 0.0) + bonus
            if worst <= target_log2_size + 1e-9:
                break
            if not weight or len(sliced) >= max_sliced:
                raise ValueError("Slicing budget exhausted before reaching the target size.")
            w = max(sorted(weight), key=lambda x: weight[x])
            sliced.append(w)
            for ws in sets:
                ws.discard(w)
        return self.with_sliced(sliced)

    # -- serialisation -----------------------------------------------------------
    def to_json(self) -> dict[str, Any]:
        return {
            "network": self.network.to_json(),
            "ssa_path": [list(p) for p in self.ssa_path],
            "sliced_wires": list(self.sliced_wires),
        }

    def dumps(self) -> str:
        return json.dumps(self.to_json(), separators=(",", ":"))

    @classmethod
    def from_json(cls, d: dict[str, Any]) -> "ContractionTree":
        return cls(Network.from_json(d["network"]), d["ssa_path"], d.get("sliced_wires", ()))

    @classmethod
    def loads(cls, s: str) -> "ContractionTree":
        return cls.from_json(json.loads(s))
