#!/usr/bin/env python
"""
TEST INFRASTRUCTURE: finds one satisfying assignment of the wide BASELINE instances (random 3-SAT n=120 / n=200,
seed 0) with a small DPLL (unit propagation, most-constrained-variable branching) and writes it to
tests/golden/models.json.  A witness is self-certifying (tests check it against every clause); it is used to pin
the sliced wires of the sampled sub-cubes consistently, so that the sub-cube's exact answer is known to be
SAT / count >= 1 instead of a trivial 0 (tools/sampled_slices.py, tests/test_gpu_wide_configs.py).

    python tools/find_models.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.setrecursionlimit(10000)


def dpll(n, clauses, budget=5_000_000):
    """Returns a model (list of 0/1) or None (UNSAT); raises if the node budget is exhausted."""
    occ = [[] for _ in range(2 * n)]          # literal index -> clause ids; literal l>0 -> 2(l-1), l<0 -> 2(-l-1)+1
    lit = lambda l: 2 * (abs(l) - 1) + (l < 0)
    for ci, c in enumerate(clauses):
        for l in c:
            occ[lit(l)].append(ci)
    val = [-1] * n
    nodes = [0]

    def value(l):
        v = val[abs(l) - 1]
        return -1 if v < 0 else int(v == (l > 0))

    def propagate(trail):
        changed = True
        while changed:
            changed = False
            for c in clauses:
                free, sat = None, False
                nfree = 0
                for l in c:
                    v = value(l)
                    if v == 1:
                        sat = True
                        break
                    if v < 0:
                        nfree += 1
                        free = l
                if sat:
                    continue
                if nfree == 0:
                    return False
                if nfree == 1:
                    val[abs(free) - 1] = int(free > 0)
                    trail.append(abs(free) - 1)
                    changed = True
        return True

    def solve():
        nodes[0] += 1
        if nodes[0] > budget:
            raise RuntimeError("node budget exhausted")
        trail = []
        if not propagate(trail):
            for v in trail:
                val[v] = -1
            return False
        # branch on the variable in the most short open clauses
        score = {}
        for c in clauses:
            if any(value(l) == 1 for l in c):
                continue
            fr = [l for l in c if value(l) < 0]
            for l in fr:
                score[abs(l) - 1] = score.get(abs(l) - 1, 0) + (4 if len(fr) == 2 else 1)
        if not score:
            for i in range(n):
                if val[i] < 0:
                    val[i] = 0
            return True
        v = max(sorted(score), key=lambda x: score[x])
        for b in (1, 0):
            val[v] = b
            if solve():
                return True
            val[v] = -1
        for u in trail:
            val[u] = -1
        return False

    return list(val) if solve() else None


def main():
    from tensorsat_b200.lib.sat import CNFInstance

    out = {}
    for name, n, m, seeds in (("c2", 120, 511, range(0, 8)), ("c4", 200, 852, range(0, 8))):
        for seed in seeds:
            cnf = CNFInstance.random(3, n, m, rng=seed)
            model = dpll(n, [tuple(c) for c in cnf.clauses])
            print(name, "seed", seed, "SAT" if model else "UNSAT", flush=True)
            out[f"{name}_seed{seed}"] = {"k": 3, "n": n, "m": m, "seed": seed, "model": model}
            if seed == 0 and model:
                assert all(any((model[abs(l) - 1] == 1) == (l > 0) for l in c) for c in cnf.clauses)
            if model and seed >= 0 and any(v["model"] for k, v in out.items() if k.startswith(name)) and seed >= 0:
                if out.get(f"{name}_seed0", {}).get("model") or seed > 0:
                    break
    with open(os.path.join(ROOT, "tests", "golden", "models.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("written tests/golden/models.json")


if __name__ == "__main__":
    main()
