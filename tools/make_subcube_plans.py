#!/usr/bin/env python
"""
Writes bench_plans/subcube_{c2,c4,c5}.json (tensorsat_b200/planner/subcube.py): network, leaves, elimination-order
tree sliced to 2^30-entry intermediates, and the witness value of every sliced wire.

    python tools/make_subcube_plans.py [c2 c4 c5]

Witnesses: tests/golden/models.json (tools/find_models.py) for the CNFs, the factors 65521 x 65519 for the
multiplier.  Every other wire value follows by propagation through the gate relations
(planner.sampling.propagate_wire_values); the script asserts that ALL wires are determined and that every leaf
relation holds at the witness.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

P, Q = 65521, 65519
N_PRIME = 4294967291   # 2^32 - 5, prime: no 16 x 16-bit factorisation, every sub-cube of that instance is empty


def multiplier(nbits, N):
    from tensorsat_b200.diagrams import DiagramBuilder
    from tensorsat_b200.lib import bincirc as bc

    mult = bc.wallace_multiplier(nbits)
    b = DiagramBuilder()
    ins = [(bc.bit_unk @ b)[0] for _ in range(2 * nbits)]
    outs = mult @ b[ins]
    for i, w in enumerate(outs):
        (bc.bit_1 if (N >> i) & 1 else bc.bit_0) @ b[w]
    return b.diagram().flatten(), ins


def main():
    from tensorsat_b200.lib.sat import CNFInstance
    from tensorsat_b200.planner import Network, propagate_wire_values
    from tensorsat_b200.planner.search import find_tree_elimination
    from tensorsat_b200.planner.subcube import SubcubeFile, path_of

    which = sys.argv[1:] or ["c2", "c4", "c5"]
    with open(os.path.join(ROOT, "tests", "golden", "models.json")) as f:
        models = json.load(f)
    for config in which:
        t0 = time.time()
        extra = {}
        if config in ("c2", "c4"):
            n, m = (120, 511) if config == "c2" else (200, 852)
            cnf = CNFInstance.random(3, n, m, rng=0)
            d = (cnf.inputs(None) >> cnf.diagram()).flatten()
            model = models[f"{config}_seed0"]["model"]
            assert all(any((model[abs(l) - 1] == 1) == (l > 0) for l in c) for c in cnf.clauses)
            known = {i: v for i, v in enumerate(model)}
            desc = f"random 3-SAT n={n} m={m} seed=0 (ratio 4.26), all inputs unknown, bintree network"
            mode, n_vars = ("count" if config == "c2" else "sat"), n
            extra["model"] = model
        else:
            d, ins = multiplier(16, P * Q)
            known = {w: (P >> i) & 1 for i, w in enumerate(ins[:16])}
            known.update({w: (Q >> i) & 1 for i, w in enumerate(ins[16:])})
            desc = f"wallace_multiplier(16) with the product fixed to {P}*{Q} = {P * Q}"
            mode, n_vars = "sat", 32
            d2, _ = multiplier(16, N_PRIME)
            assert d2.wiring == d.wiring
            extra["perturbed_product"] = N_PRIME
            extra["perturbed_leaves"] = [[list(b.tensor.shape), [int(x) for x in b.tensor.reshape(-1)]] for b in d2.boxes]
            extra["factors"] = [P, Q]
        net = Network.from_diagram(d)
        leaves = [np.asarray(b.tensor, dtype=np.uint8) for b in d.boxes]
        vals = propagate_wire_values(net, leaves, known)
        assert len(vals) == net.num_wires, (len(vals), net.num_wires)
        for ws, t in zip(net.inputs, leaves):
            assert t[tuple(vals[w] for w in ws)] == 1
        tree = find_tree_elimination(net, trials=1, seed=0)
        full = tree.slice_to_fast(30)
        f = SubcubeFile(config, desc, mode, n_vars, full, leaves, {w: vals[w] for w in full.sliced_wires}, extra)
        extra["tree_log2_width"] = tree.width()
        extra["n_sliced"] = len(full.sliced_wires)
        with open(path_of(config), "w") as fh:
            json.dump(f.to_json(), fh, separators=(",", ":"))
        print(config, "leaves", net.num_leaves, "wires", net.num_wires, "width", tree.width(), "sliced", len(full.sliced_wires),
              "file", os.path.getsize(path_of(config)) >> 10, "KiB", f"{time.time() - t0:.1f} s", flush=True)


if __name__ == "__main__":
    main()
