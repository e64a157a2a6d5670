"""
Satisfying-assignment extraction by self-reduction on ONE device plan.

The reference does this in its notebooks (notebooks/02-random-ksat.ipynb:707-722,
01-introduction.ipynb:1824-1849): fix the variables one by one, trying 0 before 1, and keep a value
when the formula stays satisfiable with the remaining variables unknown.  Every candidate is the same
network with different *input leaves* (``CNFInstance.inputs``, tensorsat/lib/sat.py:165-206), so the
contraction tree is planned once (``CNFInstance.contraction_wiring``, sat.py:231-246) and each decision is
one ``TreeContraction.contract_from_host`` call on that plan: leaves packed into one pinned buffer, one
H2D copy, one verdict byte back (plan re-use).
"""

from __future__ import annotations

from collections.abc import Callable

import numpy as np

from ..contractions.tree import TreeContraction
from .sat import CNFInstance

_BIT = {"0": np.array([1, 0], dtype=np.uint8), "1": np.array([0, 1], dtype=np.uint8),
        "?": np.array([1, 1], dtype=np.uint8)}


class AssignmentSearch:
    """Decides partial assignments of a CNF formula on one reusable plan and extracts models."""

    def __init__(self, cnf: CNFInstance, contraction: TreeContraction, *, propagate: bool = True) -> None:
        """``propagate``: let unit propagation (host, ``tsb_network_propagate``) answer the decisions it can — a
        candidate it refutes is unsatisfiable, and when the other value of the variable is refuted under a
        satisfiable prefix the candidate is satisfiable — so that only the undetermined candidates cost a
        contraction.  The decisions and their order are those of the reference either way."""
        if contraction.wiring != cnf.contraction_wiring():
            raise ValueError("The contraction was planned for a different wiring.")
        self._cnf, self._contraction = cnf, contraction
        self._up = None
        if propagate:
            from ..frontend import NativeNetwork

            self._up = NativeNetwork.from_clauses(cnf.num_vars, cnf.clauses)
        self.implied = 0
        n = cnf.num_vars
        zeros = [b.tensor for b in (cnf.inputs("0" * n) >> cnf.diagram()).flatten().boxes]
        ones = [b.tensor for b in (cnf.inputs("1" * n) >> cnf.diagram()).flatten().boxes]
        # the input boxes are the leaves that differ between the two; they come in variable order
        self._slots = [i for i, (a, b) in enumerate(zip(zeros, ones)) if not np.array_equal(a, b)]
        if len(self._slots) != n:
            raise ValueError("Could not locate the input leaves of the formula's network.")
        self._leaves = list(zeros)
        self.decisions = 0

    def satisfiable(self, spec: str | None = None) -> bool:
        """Is there a model that extends the partial assignment ``spec`` ('0' / '1' / '?' per variable)?"""
        n = self._cnf.num_vars
        spec = "?" * n if spec is None else spec
        if len(spec) != n or set(spec) - set("01?"):
            raise ValueError(f"Expected {n} characters out of '0', '1', '?'.")
        for slot, ch in zip(self._slots, spec):
            self._leaves[slot] = _BIT[ch]
        self.decisions += 1
        return bool(np.all(self._contraction.contract_from_host(self._leaves) == 1))

    def first_model(self, on_decision: Callable[[str, bool], None] | None = None) -> str | None:
        """The lexicographically first model, or ``None`` if the formula is unsatisfiable: at most 2n + 1
        decisions, each one execution of the plan."""
        n = self._cnf.num_vars

        def report(spec: str, ok: bool) -> bool:
            if on_decision is not None:
                on_decision(spec, ok)
            return ok

        def decide(spec: str, other: str | None = None) -> bool:
            """``other``: the same prefix with the other value of the variable, given when the prefix is known to
            be satisfiable (then exactly one refutation settles both candidates)."""
            if self._up is not None:
                if self._up.propagate(spec)[0]:
                    self.implied += 1
                    return report(spec, False)
                if other is not None and self._up.propagate(other)[0]:
                    self.implied += 1
                    return report(spec, True)
            return report(spec, self.satisfiable(spec))

        if not decide("?" * n):
            return None
        prefix = ""
        for _ in range(n):
            rest = "?" * (n - len(prefix) - 1)
            zero, one = prefix + "0" + rest, prefix + "1" + rest
            if decide(zero, one):
                prefix += "0"
            elif self._up is not None:
                # `prefix` is satisfiable and 0 is not: 1 is, without a contraction (the reference runs it)
                self.implied += 1
                report(one, True)
                prefix += "1"
            else:
                # the formula is satisfiable under `prefix`, so if 0 fails 1 must hold: the reference still
                # runs the check, and so does the pure path (it costs one execution)
                if not decide(one):
                    raise RuntimeError("Inconsistent decisions: neither value extends a satisfiable prefix.")
                prefix += "1"
        return prefix
