"""
TEST INFRASTRUCTURE (checker, never on the product path): an exact model counter that shares nothing with the
tensor-network path — DPLL with unit propagation, connected-component decomposition and component caching — used
to check #SAT results of sampled sub-cubes of instances the numpy oracle cannot contract (BASELINE config 2).
It counts assignments of ALL ``n`` variables that satisfy every clause (free variables count twice each), which is
what ``np.einsum`` over the relation network counts when only variable wires are free
(notebooks/prototypes/2025-02-13-#sat-tensor-contraction.ipynb:176,215).
"""

from __future__ import annotations

import sys
from collections.abc import Iterable, Sequence


def count_models(n: int, clauses: Iterable[Sequence[int]], fixed: dict[int, int] | None = None,
                 node_budget: int = 5_000_000) -> int:
    """Number of assignments of variables 1..n satisfying all ``clauses`` (DIMACS literals) with the variables in
    ``fixed`` ({var index 0-based: 0/1}) held at their values."""
    sys.setrecursionlimit(100000)
    fixed = dict(fixed or {})
    cls: list[tuple[int, ...]] = []
    for c in clauses:
        c = tuple(dict.fromkeys(int(l) for l in c))
        if any(-l in c for l in c):
            continue
        cls.append(c)
    nodes = [0]
    cache: dict[frozenset, int] = {}

    def simplify(cl: list[tuple[int, ...]], assign: dict[int, int]):
        """Unit propagation; returns (clauses left, assignment) or None on conflict."""
        assign = dict(assign)
        while True:
            out, unit = [], None
            for c in cl:
                sat, rest = False, []
                for l in c:
                    v = assign.get(abs(l) - 1)
                    if v is None:
                        rest.append(l)
                    elif v == (l > 0):
                        sat = True
                        break
                if sat:
                    continue
                if not rest:
                    return None
                if len(rest) == 1 and unit is None:
                    unit = rest[0]
                out.append(tuple(rest))
            if unit is None:
                return out, assign
            assign[abs(unit) - 1] = int(unit > 0)
            cl = out

    def components(cl):
        parent: dict[int, int] = {}

        def find(x):
            while parent.setdefault(x, x) != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x

        for c in cl:
            r = find(abs(c[0]))
            for l in c[1:]:
                parent[find(abs(l))] = r
        groups: dict[int, list] = {}
        for c in cl:
            groups.setdefault(find(abs(c[0])), []).append(c)
        return list(groups.values())

    def count(cl) -> int:
        """Models of the clause set over exactly the variables that occur in it."""
        if not cl:
            return 1
        key = frozenset(cl)
        if key in cache:
            return cache[key]
        nodes[0] += 1
        if nodes[0] > node_budget:
            raise RuntimeError("model counter: node budget exhausted")
        vars_here = {abs(l) for c in cl for l in c}
        score: dict[int, int] = {}
        for c in cl:
            for l in c:
                score[abs(l)] = score.get(abs(l), 0) + (8 >> min(len(c), 3))
        v = max(sorted(score), key=lambda x: score[x])
        total = 0
        for b in (0, 1):
            r = simplify(cl, {v - 1: b})
            if r is None:
                continue
            rest, assign = r
            left = {abs(l) for c in rest for l in c}
            free = len(vars_here) - len(left) - sum(1 for u in assign if u + 1 in vars_here)
            sub = 1
            for comp in components(rest):
                sub *= count(sorted(comp))
                if sub == 0:
                    break
            total += sub << free
        cache[key] = total
        return total

    r = simplify(cls, fixed)
    if r is None:
        return 0
    rest, assign = r
    left = {abs(l) for c in rest for l in c}
    free = n - len(left) - len(assign)
    total = 1
    for comp in components(rest):
        total *= count(sorted(comp))
    return total << free


def network_to_cnf(inputs: Sequence[Sequence[int]], leaves: Sequence) -> list[tuple[int, ...]]:
    """The relation network as a CNF over its WIRES (variable w+1 = wire w, binary wires only): every zero entry
    of a leaf tensor forbids that combination of its wires' values.  Models of the CNF = joint wire assignments
    at which every leaf holds, i.e. the terms of the network's einsum that are 1."""
    import numpy as np

    clauses = []
    for ws, t in zip(inputs, leaves):
        t = np.asarray(t)
        assert all(d == 2 for d in t.shape), "binary wires only"
        for idx in np.argwhere(t == 0):
            # forbid wires == idx  (a wire attached to several ports with conflicting values is never hit: skip)
            want = {}
            ok = True
            for w, v in zip(ws, idx):
                if want.setdefault(int(w), int(v)) != int(v):
                    ok = False
            if ok:
                clauses.append(tuple(-(w + 1) if v else (w + 1) for w, v in sorted(want.items())))
    return clauses
