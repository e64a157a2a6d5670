"""
CNF instances as relation networks: the benchmark inputs of the contraction hot path
(host-side counterpart of the reference's ``tensorsat/lib/sat.py:43-262``).

Two things here are part of the benchmark definition and are pinned by ``tests/golden``:
the random generator must consume the numpy RNG stream exactly like the reference
(``sat.py:63-78``), and the "bintree" network must number wires and order blocks exactly like
``sat.py:248-262``.
"""

from __future__ import annotations

import re
from collections.abc import Iterator, Sequence
from math import comb

import numpy as np

from ..diagrams import Diagram, DiagramBuilder, Wire, Wiring
from . import bincirc as bc

Clause = tuple[int, ...]
_HEADER = re.compile(r"p cnf ([0-9]+) ([0-9]+)")
_INPUT_BOXES = {"0": bc.bit_0, "1": bc.bit_1, "?": bc.bit_unk}


def _draw_clauses(k: int, n: int, rng: np.random.Generator) -> Iterator[Clause]:
    """Endless stream of random k-clauses; two RNG calls per clause, as in the reference."""
    variables, polarities = range(n), range(2)
    while True:
        chosen = rng.choice(variables, size=k, replace=False)
        negated = rng.choice(polarities, size=k)
        yield tuple(sorted(-(int(v) + 1) if int(s) else int(v) + 1 for v, s in zip(chosen, negated)))


def parse_dimacs(text: str) -> tuple[int, tuple[Clause, ...]]:
    """DIMACS with the reference's conventions (``sat.py:80-100``): comment lines start with ``c``,
    one clause per line, literals separated by single spaces, the trailing ``0`` is dropped."""
    rows = [ln.strip() for ln in text.strip().split("\n")]
    rows = [ln for ln in rows if ln and not ln.startswith("c")]
    head = _HEADER.match(rows[0]) if rows else None
    if head is None:
        raise ValueError("DIMACS code must start with 'p cnf <num vars> <num clauses>'.")
    num_vars, num_clauses = int(head.group(1)), int(head.group(2))
    clauses = tuple(tuple(int(tok) for tok in ln.split(" "))[:-1] for ln in rows[1:])
    if len(clauses) != num_clauses:
        raise ValueError("Number of clauses does not match the specified number.")
    return num_vars, clauses


def _check_literals(num_vars: int, clauses: Sequence[Sequence[int]]) -> None:
    if any(abs(lit) > num_vars for clause in clauses for lit in clause):
        raise ValueError("Clauses contain invalid variables.")


class CNFInstance:
    """A SAT instance in conjunctive normal form."""

    def __init__(self, num_vars: int, clauses: Sequence[Sequence[int]]) -> None:
        frozen = tuple(tuple(int(lit) for lit in clause) for clause in clauses)
        _check_literals(num_vars, frozen)
        self._n = int(num_vars)
        self._clauses = frozen

    # ------------------------------------------------------------------ constructors
    @classmethod
    def random(cls, k: int, n: int, m: int, *, rng: int | np.random.Generator | None = None) -> "CNFInstance":
        """``m`` distinct random ``k``-clauses over ``n`` variables (https://arxiv.org/abs/1405.3558)."""
        if k <= 0:
            raise ValueError("Clause size 'k' must be positive.")
        if n < k:
            raise ValueError(f"Number of variables 'n' must be at least {k = }.")
        if m <= 0:
            raise ValueError("Number of clauses 'm' must be positive.")
        limit = comb(n, k) * 2**k
        if m > limit:
            raise ValueError(f"Number of clauses 'm' cannot exceed {limit}")
        gen = rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)
        picked: dict[Clause, None] = {}
        for clause in _draw_clauses(k, n, gen):
            picked.setdefault(clause)
            if len(picked) == m:
                break
        return cls(n, tuple(picked))

    @classmethod
    def from_dimacs(cls, dimacs: str) -> "CNFInstance":
        if not isinstance(dimacs, str):
            raise TypeError("DIMACS code must be a string.")
        return cls(*parse_dimacs(dimacs))

    # ------------------------------------------------------------------ data
    @property
    def num_vars(self) -> int:
        return self._n

    @property
    def clauses(self) -> tuple[Clause, ...]:
        return self._clauses

    def to_dimacs(self) -> str:
        body = [" ".join(str(lit) for lit in clause) + " 0" for clause in self._clauses]
        return "\n".join([f"p cnf {self._n} {len(self._clauses)}", *body])

    # ------------------------------------------------------------------ networks
    def _input_spec(self, inputs: str | int | None) -> str:
        if inputs is None:
            return "?" * self._n
        if isinstance(inputs, int):
            if inputs < 0:
                raise ValueError("Inputs must be specified by a non-negative number.")
            spec = format(inputs, "b").rjust(self._n, "0")
        elif isinstance(inputs, str):
            spec = inputs
            if set(spec) - set("01?_"):
                raise ValueError("Input specification characters must be '0', '1', '?' or '_'.")
        else:
            raise TypeError("Input specification must be a string, an int or None.")
        if len(spec) != self._n:
            raise ValueError(
                f"Expected input specification string to be {self._n} characters long,"
                f" found {len(spec)} characters instead."
            )
        return spec

    def inputs(self, inputs: str | int | None, /, *, discard_unk: bool = False) -> Diagram:
        """One box per variable — ``'0'`` / ``'1'`` fix it, ``'?'`` leaves it unknown (``sat.py:165-206``).

        With ``discard_unk`` unknown variables get a bare wire instead of a ``bit_unk`` box.  ``'_'`` passes
        the reference's validation but its builder branch is dead code there (``sat.py:202``); it is rejected.
        """
        builder = DiagramBuilder()
        for var, ch in enumerate(self._input_spec(inputs)):
            if ch == "_":
                raise ValueError("Open inputs ('_') are not supported.")
            if ch == "?" and discard_unk:
                wire = builder.add_wire(bc.bit)
            else:
                (wire,) = _INPUT_BOXES[ch] @ builder
            assert wire == var
        builder.add_outputs(builder.wires)
        return builder.diagram()

    def diagram(self, *, mode: str = "bintree") -> Diagram:
        if mode != "bintree":
            raise NotImplementedError(f"Unknown diagram mode for CNFInstance: {mode!r}")
        builder = DiagramBuilder()
        builder.add_inputs(bc.bits(self._n))
        for clause in self._clauses:
            self._add_clause(builder, clause)
        return builder.diagram()

    @staticmethod
    def _add_clause(builder: DiagramBuilder, clause: Clause) -> None:
        """NOT gates for the negative literals, a balanced tree of OR gates, ``bit_1`` on its root."""
        level: list[Wire] = []
        for lit in clause:
            level.append(lit - 1 if lit > 0 else (bc.not_ @ builder[-lit - 1])[0])
        while len(level) > 1:
            pairs = len(level) // 2
            merged = [(bc.or_ @ builder[level[2 * p], level[2 * p + 1]])[0] for p in range(pairs)]
            if len(level) % 2:
                merged[-1] = (bc.or_ @ builder[merged[-1], level[-1]])[0]
            level = merged
        bc.bit_1 @ builder[level[0]]

    def contraction_wiring(self, *, mode: str = "bintree") -> Wiring:
        """The wiring shared by ``inputs(x) >> diagram()`` for every ``x`` (``sat.py:231-246``)."""
        return (self.inputs(0) >> self.diagram(mode=mode)).flatten().wiring

    def __repr__(self) -> str:
        sizes = sorted({len(clause) for clause in self._clauses})
        k = f"k={sizes[0]}" if len(sizes) == 1 else f"k in [{sizes[0]}..{sizes[-1]}]"
        return f"<CNFInstance: {self._n} vars, {len(self._clauses)} clauses, {k}>"
