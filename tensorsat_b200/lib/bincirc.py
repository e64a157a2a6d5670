"""
Binary-circuit gates and arithmetic circuits, stated as in the reference
(tensorsat/lib/bincirc.py:28-177).  They only define the *inputs* of the contraction
hot path (leaf tensors + wiring); the leaves are 2-, 4- or 8-byte 0/1 tensors that are
uploaded once per contraction.
"""

from __future__ import annotations

from itertools import product
from typing import Final

from ..diagrams import Diagram, DiagramBuilder, FinSet, Wire
from ..lang.fin_rel import FinRel

bit: Final[FinSet] = FinSet(2)
"""The set {0, 1} of binary values (bincirc.py:28)."""


def bits(n: int) -> tuple[FinSet, ...]:
    """Shape consisting of the given number of bits (bincirc.py:32-36)."""
    if n < 0:
        raise ValueError("Number of bits must be non-negative.")
    return (bit,) * n


bits2: Final = bits(2)

# Gate relations as function graphs, input ports first (bincirc.py:42-73).
not_: Final[FinRel] = FinRel.from_callable(bit, bit, lambda b: 1 - b, name="not_")
and_: Final[FinRel] = FinRel.from_callable(bits2, bit, lambda a, b: a & b, name="and_")
or_: Final[FinRel] = FinRel.from_callable(bits2, bit, lambda a, b: a | b, name="or_")
impl_: Final[FinRel] = FinRel.from_callable(bits2, bit, lambda a, b: (1 - a) | b, name="impl_")
biimpl_: Final[FinRel] = FinRel.from_callable(
    bits2, bit, lambda a, b: (a & b) | ((1 - a) & (1 - b)), name="biimpl_"
)
xor_: Final[FinRel] = FinRel.from_callable(bits2, bit, lambda a, b: a ^ b, name="xor_")
bit_0: Final[FinRel] = FinRel.singleton(bit, 0, name="bit_0")
bit_1: Final[FinRel] = FinRel.singleton(bit, 1, name="bit_1")
bit_unk: Final[FinRel] = FinRel.from_set(bit, {0, 1}, name="bit_unk")


@Diagram.from_recipe
def half_adder(diag: DiagramBuilder) -> None:
    """Half adder (bincirc.py:91-103)."""
    a, b = diag.add_inputs(bits(2))
    (s,) = xor_ @ diag[a, b]
    (c,) = and_ @ diag[a, b]
    diag.add_outputs([s, c])


@Diagram.from_recipe
def full_adder(diag: DiagramBuilder) -> None:
    """Full adder (bincirc.py:106-121)."""
    a, b, c_in = diag.add_inputs(bits(3))
    (x1,) = xor_ @ diag[a, b]
    (x2,) = and_ @ diag[a, b]
    (x3,) = and_ @ diag[x1, c_in]
    (s,) = xor_ @ diag[x1, c_in]
    (c_out,) = or_ @ diag[x2, x3]
    diag.add_outputs([s, c_out])


@Diagram.recipe
def rc_adder(diag: DiagramBuilder, num_bits: int) -> None:
    """Ripple-carry adder (bincirc.py:124-142)."""
    if num_bits <= 0:
        raise ValueError("Number of bits must be positive.")
    a = diag.add_inputs(bits(num_bits))
    b = diag.add_inputs(bits(num_bits))
    c = diag.add_input(bit)
    for i in range(num_bits):
        s, c = full_adder @ diag[a[i], b[i], c]
        diag.add_output(s)
    diag.add_output(c)


@Diagram.recipe
def wallace_multiplier(diag: DiagramBuilder, num_bits: int) -> None:
    """Wallace-tree multiplier (bincirc.py:145-177)."""
    if num_bits <= 0:
        raise ValueError("Number of bits must be positive.")
    a = diag.add_inputs(bits(num_bits))
    b = diag.add_inputs(bits(num_bits))
    layer: dict[int, list[Wire]] = {}
    for i, j in product(range(num_bits), repeat=2):
        (_out,) = and_ @ diag[a[i], b[j]]
        layer.setdefault(i + j, []).append(_out)
    while any(len(wires) > 1 for wires in layer.values()):
        new_layer: dict[int, list[Wire]] = {}
        for weight, wires in layer.items():
            num_fulladd, _r = divmod(len(wires), 3)
            for idx in range(num_fulladd):
                s, c_out = full_adder @ diag[wires[3 * idx : 3 * idx + 3]]
                new_layer.setdefault(weight, []).append(s)
                new_layer.setdefault(weight + 1, []).append(c_out)
            if _r == 2:
                s, c_out = half_adder @ diag[wires[-2:]]
                new_layer.setdefault(weight, []).append(s)
                new_layer.setdefault(weight + 1, []).append(c_out)
            elif _r == 1:
                new_layer.setdefault(weight, []).append(wires[-1])
        layer = new_layer
    assert all(len(wires) == 1 for wires in layer.values())
    diag.add_outputs(wires[0] for wires in layer.values())
