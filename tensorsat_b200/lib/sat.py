"""
CNF instances and their relation networks, stated as in the reference
(tensorsat/lib/sat.py:43-262).  The generator must reproduce the reference's RNG stream
exactly, because the benchmark instances are defined by ``(k, n, m, seed)``.
"""

from __future__ import annotations

import re
from collections.abc import Sequence
from math import comb

import numpy as np

from ..diagrams import Diagram, DiagramBuilder, Wire, Wiring
from .bincirc import bit, bit_0, bit_1, bit_unk, bits, not_, or_

Clause = tuple[int, ...]


class CNFInstance:
    """A SAT instance in CNF form (tensorsat/lib/sat.py:40-275)."""

    __slots__ = ("_num_vars", "_clauses")

    def __init__(self, num_vars: int, clauses: Sequence[Sequence[int]]) -> None:
        clauses = tuple(tuple(int(x) for x in c) for c in clauses)
        if num_vars < max(abs(lit) for clause in clauses for lit in clause):
            raise ValueError("Clauses contain invalid variables.")
        self._num_vars = int(num_vars)
        self._clauses = clauses

    @classmethod
    def random(
        cls, k: int, n: int, m: int, *, rng: int | np.random.Generator | None = None
    ) -> "CNFInstance":
        """Random k-SAT, n variables, m distinct clauses (sat.py:43-78; same RNG calls)."""
        if k <= 0:
            raise ValueError("Clause size 'k' must be positive.")
        if n < k:
            raise ValueError(f"Number of variables 'n' must be at least {k = }.")
        if m <= 0:
            raise ValueError("Number of clauses 'm' must be positive.")
        if m > comb(n, k) * 2**k:
            raise ValueError(f"Number of clauses 'm' cannot exceed {comb(n, k)*2**k = }")
        if not isinstance(rng, np.random.Generator):
            rng = np.random.default_rng(rng)
        clauses: list[Clause] = []
        seen: set[Clause] = set()
        while len(clauses) < m:
            vs = [int(v) for v in rng.choice(range(n), size=k, replace=False)]
            signs = [int(s) for s in rng.choice(range(2), size=k)]
            clause = tuple(sorted(v + 1 if s == 0 else -v - 1 for v, s in zip(vs, signs)))
            if clause not in seen:
                clauses.append(clause)
                seen.add(clause)
        return cls(n, clauses)

    @classmethod
    def from_dimacs(cls, dimacs: str) -> "CNFInstance":
        """DIMACS parser with the reference's conventions (sat.py:80-100): one clause per
        line, single-space separated, trailing ``0`` dropped, lines starting with ``c``
        skipped."""
        if not isinstance(dimacs, str):
            raise TypeError("DIMACS code must be a string.")
        lines = [
            stripped
            for line in dimacs.strip().split("\n")
            if (stripped := line.strip()) and not stripped.startswith("c")
        ]
        start = re.compile(r"p cnf ([0-9]+) ([0-9]+)").match(lines[0]) if lines else None
        if not start:
            raise ValueError("DIMACS code must start with 'p cnf <num vars> <num clauses>'.")
        num_vars, num_clauses = map(int, start.groups())
        clauses = tuple(tuple(map(int, line.split(" ")))[:-1] for line in lines[1:])
        if len(clauses) != num_clauses:
            raise ValueError("Number of clauses does not match the specified number.")
        if num_vars < max(abs(lit) for clause in clauses for lit in clause):
            raise ValueError("Clauses contain invalid variables.")
        return cls(num_vars, clauses)

    @property
    def num_vars(self) -> int:
        return self._num_vars

    @property
    def clauses(self) -> tuple[Clause, ...]:
        return self._clauses

    def to_dimacs(self) -> str:
        lines = [
            f"p cnf {self.num_vars} {len(self.clauses)}",
            *(f"{' '.join(map(str, clause))} 0" for clause in self.clauses),
        ]
        return "\n".join(lines)

    def _normalize_inputs(self, inputs: str | int | None) -> str:
        n = self.num_vars
        if isinstance(inputs, int):
            if inputs < 0:
                raise ValueError("Inputs must be specified by a non-negative number.")
            inputs = f"{inputs:0>{n}b}"
        elif inputs is None:
            inputs = "?" * n
        else:
            if not isinstance(inputs, str):
                raise TypeError("Input specification must be a string, an int or None.")
            if not all(b in "01?_" for b in inputs):
                raise ValueError("Input specification characters must be '0', '1', '?' or '_'.")
        if len(inputs) != n:
            raise ValueError(
                f"Expected input specification string to be {n} characters long,"
                f" found {len(inputs)} characters instead."
            )
        return inputs

    def inputs(self, inputs: str | int | None, /, *, discard_unk: bool = False) -> Diagram:
        """Diagram fixing each variable to 0 / 1 / unknown (sat.py:165-206).

        ``'_'`` (leave the input open) is accepted by the reference's validation but its
        builder branch is dead code (``case None`` never matches a character, sat.py:202)
        and trips the ``assert w == i`` that follows; here it raises ``ValueError``.
        """
        inputs = self._normalize_inputs(inputs)
        builder = DiagramBuilder()
        for i, b in enumerate(inputs):
            w: Wire
            if b == "0":
                (w,) = bit_0 @ builder
            elif b == "1":
                (w,) = bit_1 @ builder
            elif b == "?":
                if discard_unk:
                    w = builder.add_wire(bit)
                else:
                    (w,) = bit_unk @ builder
            else:
                raise ValueError("Open inputs ('_') are not supported.")
            assert w == i
        builder.add_outputs(builder.wires)
        return builder.diagram()

    def diagram(self, *, mode: str = "bintree") -> Diagram:
        if mode == "bintree":
            return self._build_diagram_bintree()
        raise NotImplementedError(f"Unknown diagram mode for CNFInstance: {mode!r}")

    def contraction_wiring(self, *, mode: str = "bintree") -> Wiring:
        """Wiring shared by every ``inputs(...) >> diagram()`` (sat.py:231-246)."""
        return (self.inputs(0) >> self.diagram(mode=mode)).flatten().wiring

    def _build_diagram_bintree(self) -> Diagram:
        """One balanced OR tree per clause, ``bit_1`` on its output (sat.py:248-262)."""
        builder = DiagramBuilder()
        builder.add_inputs(bits(self.num_vars))
        for clause in self.clauses:
            layer = [x - 1 if x > 0 else (not_ @ builder[-x - 1])[0] for x in clause]
            while (n := len(layer)) > 1:
                new_layer = [(or_ @ builder[layer[2 * i : 2 * i + 2]])[0] for i in range(n // 2)]
                if n % 2 == 1:
                    (new_layer[-1],) = or_ @ builder[new_layer[-1], layer[-1]]
                layer = new_layer
            assert len(layer) == 1
            bit_1 @ builder[layer[0]]
        return builder.diagram()

    def __repr__(self) -> str:
        attrs = [f"{self.num_vars} vars", f"{len(self.clauses)} clauses"]
        ks = {len(clause) for clause in self.clauses}
        attrs.append(f"k={ks.pop()}" if len(ks) == 1 else f"k in [{min(ks)}..{max(ks)}]")
        return f"<CNFInstance: {', '.join(attrs)}>"
