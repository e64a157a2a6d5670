"""
Gate relations and arithmetic circuits that define the benchmark networks of the contraction hot
path (the host-side counterpart of the reference's ``tensorsat/lib/bincirc.py:28-177``).

Only the *inputs* of the executor are produced here: 2-, 4- and 8-byte 0/1 leaf tensors and the
wiring between them.  Block and wire creation order follows the reference exactly, because the
flat wiring (wire numbers, slot order) is part of the benchmark definition; the wiring digests of
``tests/golden/instances.json`` pin it.
"""

from __future__ import annotations

from typing import Callable, Final

from ..diagrams import Diagram, DiagramBuilder, FinSet, Wire
from ..lang.fin_rel import FinRel

bit: Final[FinSet] = FinSet(2)
"""The two-element set of binary values."""


def bits(n: int) -> tuple[FinSet, ...]:
    """Shape made of ``n`` bits."""
    if n < 0:
        raise ValueError("Number of bits must be non-negative.")
    return tuple(bit for _ in range(n))


bits2: Final = bits(2)


def _gate(name: str, arity: int, fn: Callable[..., int]) -> FinRel:
    """Graph of a boolean function: input ports first, the output port last."""
    return FinRel.from_callable(bits(arity), bit, fn, name=name)


_TRUTH: Final[dict[str, tuple[int, Callable[..., int]]]] = {
    "not_": (1, lambda a: a ^ 1),
    "and_": (2, lambda a, b: a & b),
    "or_": (2, lambda a, b: a | b),
    "impl_": (2, lambda a, b: (a ^ 1) | b),
    "biimpl_": (2, lambda a, b: (a ^ b) ^ 1),
    "xor_": (2, lambda a, b: a ^ b),
}

not_: Final[FinRel] = _gate("not_", *_TRUTH["not_"])
and_: Final[FinRel] = _gate("and_", *_TRUTH["and_"])
or_: Final[FinRel] = _gate("or_", *_TRUTH["or_"])
impl_: Final[FinRel] = _gate("impl_", *_TRUTH["impl_"])
biimpl_: Final[FinRel] = _gate("biimpl_", *_TRUTH["biimpl_"])
xor_: Final[FinRel] = _gate("xor_", *_TRUTH["xor_"])

bit_0: Final[FinRel] = FinRel.singleton(bit, 0, name="bit_0")
bit_1: Final[FinRel] = FinRel.singleton(bit, 1, name="bit_1")
bit_unk: Final[FinRel] = FinRel.from_set(bit, (0, 1), name="bit_unk")
"""Constant 0, constant 1 and the unconstrained bit {0, 1}."""


def _half_adder(circ: DiagramBuilder) -> None:
    x, y = circ.add_inputs(bits2)
    total = (xor_ @ circ[x, y])[0]
    carry = (and_ @ circ[x, y])[0]
    circ.add_outputs((total, carry))


def _full_adder(circ: DiagramBuilder) -> None:
    x, y, carry_in = circ.add_inputs(bits(3))
    partial = (xor_ @ circ[x, y])[0]
    gen = (and_ @ circ[x, y])[0]
    prop = (and_ @ circ[partial, carry_in])[0]
    total = (xor_ @ circ[partial, carry_in])[0]
    carry_out = (or_ @ circ[gen, prop])[0]
    circ.add_outputs((total, carry_out))


half_adder: Final[Diagram] = Diagram.from_recipe(_half_adder)
"""(a, b) -> (sum, carry)."""

full_adder: Final[Diagram] = Diagram.from_recipe(_full_adder)
"""(a, b, carry_in) -> (sum, carry_out)."""


@Diagram.recipe
def rc_adder(circ: DiagramBuilder, num_bits: int) -> None:
    """Ripple-carry adder of two ``num_bits``-bit numbers with carry in / carry out."""
    if num_bits <= 0:
        raise ValueError("Number of bits must be positive.")
    xs = circ.add_inputs(bits(num_bits))
    ys = circ.add_inputs(bits(num_bits))
    carry = circ.add_input(bit)
    for x, y in zip(xs, ys):
        total, carry = full_adder @ circ[x, y, carry]
        circ.add_output(total)
    circ.add_output(carry)


def _compress_column(circ: DiagramBuilder, column: list[Wire]) -> tuple[list[Wire], list[Wire]]:
    """One Wallace step on the wires of one weight: full adders on triples, a half adder on a remaining
    pair, a remaining single wire passes through.  Returns (wires kept at this weight, carries)."""
    kept: list[Wire] = []
    carries: list[Wire] = []
    triples, rest = divmod(len(column), 3)
    for t in range(triples):
        s, c = full_adder @ circ[column[3 * t : 3 * t + 3]]
        kept.append(s)
        carries.append(c)
    if rest == 2:
        s, c = half_adder @ circ[column[-2:]]
        kept.append(s)
        carries.append(c)
    elif rest == 1:
        kept.append(column[-1])
    return kept, carries


@Diagram.recipe
def wallace_multiplier(circ: DiagramBuilder, num_bits: int) -> None:
    """Wallace-tree multiplier: ``num_bits`` + ``num_bits`` input ports, then the product bits by weight."""
    if num_bits <= 0:
        raise ValueError("Number of bits must be positive.")
    xs = circ.add_inputs(bits(num_bits))
    ys = circ.add_inputs(bits(num_bits))
    columns: list[list[Wire]] = [[] for _ in range(2 * num_bits - 1)]
    for i, x in enumerate(xs):
        for j, y in enumerate(ys):
            columns[i + j].append((and_ @ circ[x, y])[0])
    while max(len(col) for col in columns) > 1:
        nxt: list[list[Wire]] = [[] for _ in range(len(columns) + 1)]
        for weight, column in enumerate(columns):
            kept, carries = _compress_column(circ, column)
            nxt[weight].extend(kept)
            nxt[weight + 1].extend(carries)
        while nxt and not nxt[-1]:
            nxt.pop()
        columns = nxt
    circ.add_outputs(col[0] for col in columns)
