"""
Generates the golden fixtures in this directory by running the reference's OWN code.

Run in the build container only (the reference lives at /root/reference there and is
absent on the GPU box):

    python tests/golden/make_golden.py

The reference is imported through ``oracle/ref_shims/load_reference.py`` (stand-ins for
its missing pure-Python dependencies; nothing of the reference is copied).  What is
recorded:

* ``ops.json``      — ``FinRel.contract2 / rewire / spider / stack / __getitem__ / __add__``
                      and ``tensorsat._autoray.einsum / transpose / stack`` outputs on seeded
                      random 0/1 tensors, including repeated / dropped output wires,
                      non-binary dimensions and a step that hits the mod-256 clamp quirk.
* ``instances.json``— ``CNFInstance.random`` clauses, flat wirings of ``inputs >> diagram``
                      (hashed + sizes), Wallace / ripple-carry wirings, ``from_dimacs`` parses.
* ``networks.json`` — whole-network results obtained by stepping the reference's
                      ``_autoray.einsum`` along a recorded SSA path: verdicts of the kissat
                      regression CNFs used by examples/test_sat_solver.py, small random
                      k-SAT solution tensors, circuit truth tables, ``SimpleContraction`` runs.
* ``kissat/*.cnf``  — the DIMACS inputs of those tests (kissat's own regression files,
                      copied verbatim as data so that the GPU box can run them).
"""

from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shims"))

import load_reference  # noqa: E402

load_reference.load()

from tensorsat.contractions.simple import SimpleContraction  # noqa: E402
from tensorsat.diagrams import Diagram  # noqa: E402
from tensorsat.lang.fin_rel import FinRel, FinSet  # noqa: E402
from tensorsat.lib import bincirc as rb  # noqa: E402
from tensorsat.lib.sat import CNFInstance  # noqa: E402
import tensorsat._autoray as ref_autoray  # noqa: E402

from tensorsat_b200.planner import ContractionTree, Network  # noqa: E402
from tensorsat_b200.planner.search import find_tree_elimination  # noqa: E402

SAT_TESTS = {  # examples/test_sat_solver.py:22-190 (name -> expected verdict)
    "def1": False, "congr1": False, "congr2": False, "congr3": False, "congr4": True,
    "congr5": False, "congr6": False, "congr7": True, "xor0": True, "xor1": True,
    "xor2": True, "xor3": True, "xor4": True, "xor5": False, "xor6": True, "ph2": False,
    "full2": False, "full3": False, "full4": False, "strash1": False, "strash3": True,
    "tieshirt": True, "prime4": True, "eq1": True, "eq2": False, "probe1": True,
    "ite12": True, "ite13": True, "ite14": True, "ite15": True, "ite16": True,
    "ite17": True, "ite18": True, "diamond1": False, "diamond2": False, "diamond3": False,
    "unit1": False, "unit2": False, "unit3": False, "prime121": True,
}
PARSER_TESTS = ["unit1", "unit2", "unit3", "unit4", "unit5", "bin1"]  # examples/test_dimacs_parser.py


def wiring_digest(wiring) -> str:
    blob = json.dumps([list(map(list, wiring.slot_wires_list)), list(wiring.out_wires)]).encode()
    return hashlib.sha256(blob).hexdigest()


def symbols(n: int) -> list[str]:
    base = [chr(c) for c in range(ord("a"), ord("z") + 1)] + [chr(c) for c in range(ord("A"), ord("Z") + 1)]
    return base + [chr(0x100 + i) for i in range(max(0, n - 52))]


def reference_execute(diagram: Diagram, tree: ContractionTree) -> FinRel:
    """Step the reference's autoray ``einsum`` (tensorsat/_autoray.py:31-62) along the tree."""
    boxes = dict(enumerate(diagram.boxes))
    wires = {i: tuple(ws) for i, ws in enumerate(diagram.wiring.slot_wires_list)}
    out = list(diagram.wiring.out_wires)
    L = len(boxes)
    for s, st in enumerate(tree.steps):
        la, lb = wires.pop(st.lhs), wires.pop(st.rhs)
        a, b = boxes.pop(st.lhs), boxes.pop(st.rhs)
        ow = [w for w in dict.fromkeys(la + lb) if w in st.out_wires]
        if s == len(tree.steps) - 1:
            ow = out
        labs = list(dict.fromkeys(la + lb))
        sym = dict(zip(labs, symbols(len(labs))))
        eq = "".join(sym[w] for w in la) + "," + "".join(sym[w] for w in lb) + "->" + "".join(sym[w] for w in ow)
        boxes[L + s] = ref_autoray.einsum(eq, a, b)
        wires[L + s] = tuple(ow)
    (res,) = boxes.values()
    return res


def gen_ops(rng: np.random.Generator) -> list[dict]:
    cases: list[dict] = []

    def rand_rel(shape):
        return FinRel((rng.random(shape) < 0.55).astype(np.uint8))

    # contract2: random wire patterns over binary and mixed dims
    for trial in range(60):
        nl, nr = int(rng.integers(0, 6)), int(rng.integers(0, 6))
        dims_pool = [2, 2, 2, 2, 3] if trial % 3 == 0 else [2]
        n_wires = max(1, nl + nr - int(rng.integers(0, min(nl, nr) + 1)))
        wdim = [int(rng.choice(dims_pool)) for _ in range(n_wires)]
        lw = [int(w) for w in rng.choice(n_wires, size=nl, replace=False)] if nl <= n_wires else list(range(nl))
        rw = [int(w) for w in rng.choice(n_wires, size=nr, replace=False)] if nr <= n_wires else list(range(nr))
        allw = sorted(set(lw) | set(rw))
        kind = trial % 4
        if kind == 0 or not allw:
            ow = None
        elif kind == 1:
            k = int(rng.integers(0, len(allw) + 1))
            ow = [int(w) for w in rng.permutation(allw)[:k]]
        elif kind == 2:  # repeated output wires
            k = int(rng.integers(1, len(allw) + 2))
            ow = [int(w) for w in rng.choice(allw, size=k, replace=True)]
        else:
            ow = [int(w) for w in rng.permutation(allw)]
        lhs = rand_rel([wdim[w] for w in lw])
        rhs = rand_rel([wdim[w] for w in rw])
        case = {
            "op": "contract2", "lhs": lhs.tensor.tolist(), "lhs_wires": lw,
            "rhs": rhs.tensor.tolist(), "rhs_wires": rw, "out_wires": ow,
            "lhs_shape": list(lhs.tensor.shape), "rhs_shape": list(rhs.tensor.shape),
        }
        try:
            res = FinRel.contract2(lhs, lw, rhs, rw, ow)
            case |= {"res": res.tensor.tolist(), "res_shape": list(res.tensor.shape)}
        except ValueError as e:  # reference rewire_array cannot permute surviving ports
            case |= {"raises": "ValueError", "message": str(e)[:80]}
        cases.append(case)
    # a step that sums exactly 256 ones: the reference clamps it to 0 (np.sign dtype quirk)
    ones8 = FinRel(np.ones((2,) * 8, dtype=np.uint8))
    res = FinRel.contract2(ones8, list(range(8)), ones8, list(range(8)), [])
    cases.append({"op": "contract2_quirk", "k": 8, "res": res.tensor.tolist()})
    ones9 = FinRel(np.ones((2,) * 9, dtype=np.uint8))
    res = FinRel.contract2(ones9, list(range(9)), ones9, list(range(9)), [8])
    cases.append({"op": "contract2_quirk", "k": 9, "res": res.tensor.tolist()})
    # rewire: permutations, drops, duplications
    for trial in range(40):
        n = int(rng.integers(0, 6))
        dims = [int(rng.choice([2, 2, 3])) for _ in range(n)]
        t = rand_rel(dims)
        kind = trial % 3
        if n == 0:
            ports = []
        elif kind == 0:
            ports = [int(p) for p in rng.permutation(n)]
        elif kind == 1:
            ports = [int(p) for p in rng.permutation(n)[: int(rng.integers(0, n + 1))]]
        else:
            ports = [int(p) for p in rng.choice(n, size=int(rng.integers(1, n + 3)), replace=True)]
        case = {"op": "rewire", "tensor": t.tensor.tolist(), "shape": dims, "ports": ports}
        try:
            res = t.rewire(ports)
            case |= {"res": np.asarray(res.tensor).tolist(), "res_shape": list(res.tensor.shape)}
        except ValueError as e:  # reference rewire_array cannot permute surviving ports
            case |= {"raises": "ValueError", "message": str(e)[:80]}
        cases.append(case)
    # spider / scalar
    for size, n in [(2, 1), (2, 2), (2, 3), (2, 5), (3, 2), (3, 3), (4, 2), (1, 3)]:
        cases.append({"op": "spider", "size": size, "num_ports": n,
                      "res": FinRel.spider(FinSet(size), n).tensor.tolist()})
    for v in (False, True):
        cases.append({"op": "scalar", "value": v, "res": FinRel.scalar(v).tensor.tolist(),
                      "bool": bool(FinRel.scalar(v))})
    # getitem / union / stack
    for trial in range(20):
        n = int(rng.integers(1, 5))
        dims = [int(rng.choice([2, 2, 3])) for _ in range(n)]
        a, b = rand_rel(dims), rand_rel(dims)
        idx = [int(rng.integers(0, d)) if rng.random() < 0.5 else None for d in dims]
        got = a[tuple(slice(None) if i is None else i for i in idx)]
        port = int(rng.integers(0, n + 1))
        cases.append({
            "op": "slice_union_stack", "a": a.tensor.tolist(), "b": b.tensor.tolist(), "shape": dims,
            "idx": idx, "getitem": np.asarray(got.tensor).tolist(), "getitem_shape": list(got.tensor.shape),
            "union": (a + b).tensor.tolist(), "port": port,
            "stack": FinRel.stack([a, b, a], port).tensor.tolist(),
            "autoray_stack": ref_autoray.stack([a, b], port).tensor.tolist(),
        })
    # autoray einsum strings / transpose
    for eq, ls, rs in [("abc,cde->abde", (2, 2, 2), (2, 2, 2)), ("ab,bc->ac", (2, 3), (3, 2)),
                       ("ab,ab->", (2, 2), (2, 2)), ("a,b->ab", (2,), (3,)),
                       ("abc,bcd->ad", (2, 2, 2), (2, 2, 2)), ("ab,bc->ca", (2, 2), (2, 2)),
                       ("abcd,ce->abde", (2, 2, 2, 2), (2, 2))]:
        a, b = rand_rel(ls), rand_rel(rs)
        cases.append({"op": "einsum", "eq": eq, "lhs": a.tensor.tolist(), "rhs": b.tensor.tolist(),
                      "res": np.asarray(ref_autoray.einsum(eq, a, b).tensor).tolist()})
    t = rand_rel((2, 3, 2))
    cases.append({"op": "transpose", "tensor": t.tensor.tolist(), "perm": [2, 0, 1],
                  "res": np.asarray(ref_autoray.transpose(t, [2, 0, 1]).tensor).tolist()})
    return cases


def gen_instances() -> dict:
    out: dict = {"random": [], "circuits": [], "dimacs": []}
    for k, n, m, seed in [(3, 4, 4, 0), (3, 3, 6, 3), (4, 10, 95, 0), (4, 20, 195, 1), (3, 10, 42, 5),
                          (3, 20, 85, 0), (3, 50, 213, 0), (3, 50, 213, 1), (3, 120, 511, 0), (3, 200, 852, 0)]:
        c = CNFInstance.random(k, n, m, rng=seed)
        d = (c.inputs(None) >> c.diagram()).flatten()
        rec = {"k": k, "n": n, "m": m, "seed": seed,
               "clauses_sha256": hashlib.sha256(json.dumps(c.clauses).encode()).hexdigest(),
               "num_wires": d.wiring.num_wires, "num_blocks": len(d.blocks),
               "wiring_sha256": wiring_digest(d.wiring),
               "diagram_num_blocks": len(c.diagram().blocks),
               "contraction_wiring_sha256": wiring_digest(c.contraction_wiring())}
        if n <= 10:
            rec["clauses"] = [list(cl) for cl in c.clauses]
            rec["slot_wires_list"] = [list(ws) for ws in d.wiring.slot_wires_list]
            rec["box_names"] = [b.name for b in d.boxes]
        out["random"].append(rec)
    for name, diag in [("half_adder", rb.half_adder), ("full_adder", rb.full_adder)] + [
        (f"rc_adder({n})", rb.rc_adder(n)) for n in (1, 2, 4)] + [
        (f"wallace_multiplier({n})", rb.wallace_multiplier(n)) for n in (1, 2, 3, 4, 8, 16)]:
        f = diag.flatten()
        rec = {"name": name, "num_wires": f.wiring.num_wires, "num_blocks": len(f.blocks),
               "num_ports": f.num_ports, "wiring_sha256": wiring_digest(f.wiring),
               "out_wires": list(f.wiring.out_wires)}
        if f.wiring.num_wires <= 60:
            rec["slot_wires_list"] = [list(ws) for ws in f.wiring.slot_wires_list]
            rec["box_names"] = [b.name for b in f.boxes]
        out["circuits"].append(rec)
    gates = {n: getattr(rb, n).tensor.tolist() for n in
             ["not_", "and_", "or_", "impl_", "biimpl_", "xor_", "bit_0", "bit_1", "bit_unk"]}
    out["gates"] = gates
    for name in sorted(set(PARSER_TESTS) | set(SAT_TESTS)):
        with open(os.path.join(load_reference.REFERENCE_ROOT, "examples/kissat/cnf", f"{name}.cnf")) as f:
            c = CNFInstance.from_dimacs(f.read())
        rec = {"name": name, "num_vars": c.num_vars, "num_clauses": len(c.clauses),
               "clauses_sha256": hashlib.sha256(json.dumps(c.clauses).encode()).hexdigest()}
        if name in PARSER_TESTS:
            rec["clauses"] = [list(cl) for cl in c.clauses]
        out["dimacs"].append(rec)
    return out


def gen_networks() -> dict:
    out: dict = {"kissat": [], "random": [], "circuits": [], "simple": []}
    os.makedirs(os.path.join(HERE, "kissat"), exist_ok=True)
    for name, expected in SAT_TESTS.items():
        src = os.path.join(load_reference.REFERENCE_ROOT, "examples/kissat/cnf", f"{name}.cnf")
        shutil.copyfile(src, os.path.join(HERE, "kissat", f"{name}.cnf"))
        with open(src) as f:
            c = CNFInstance.from_dimacs(f.read())
        d = (c.inputs(None) >> c.diagram()).flatten()
        net = Network(tuple(d.wiring.slot_wires_list), tuple(d.wiring.out_wires),
                      tuple(t.tensor_dim for t in d.wiring.wire_types))
        tree = find_tree_elimination(net, trials=4, seed=0)
        rec = {"name": name, "expected": expected, "num_vars": c.num_vars,
               "num_clauses": len(c.clauses), "log2_max_size": tree.stats().log2_max_size}
        if tree.stats().log2_max_size <= 24:
            rec["ssa_path"] = [list(p) for p in tree.ssa_path]
            res = reference_execute(d, tree)
            rec["reference_verdict"] = bool(res)
            assert bool(res) == expected, name
        out["kissat"].append(rec)
        print("kissat", name, rec.get("reference_verdict"), rec["log2_max_size"], flush=True)
    for name in PARSER_TESTS:
        src = os.path.join(load_reference.REFERENCE_ROOT, "examples/kissat/cnf", f"{name}.cnf")
        shutil.copyfile(src, os.path.join(HERE, "kissat", f"{name}.cnf"))
    # random k-SAT: solution tensors (open variable ports) and decisions with partial inputs
    for k, n, m, seed in [(3, 4, 4, 0), (3, 3, 6, 3), (4, 10, 95, 0), (3, 10, 42, 5), (3, 12, 40, 2)]:
        c = CNFInstance.random(k, n, m, rng=seed)
        d = c.diagram().flatten()
        net = Network(tuple(d.wiring.slot_wires_list), tuple(d.wiring.out_wires),
                      tuple(t.tensor_dim for t in d.wiring.wire_types))
        tree = find_tree_elimination(net, trials=4, seed=0)
        sol = reference_execute(d, tree)
        rec = {"k": k, "n": n, "m": m, "seed": seed, "ssa_path": [list(p) for p in tree.ssa_path],
               "solutions": sorted("".join(map(str, p)) for p in sol.to_set()), "decisions": []}
        for spec in [None, "0" + "?" * (n - 1), "1" + "?" * (n - 1), "01" + "?" * (n - 2),
                     "".join("01?"[(i * 7 + seed) % 3] for i in range(n))]:
            dd = (c.inputs(spec) >> c.diagram()).flatten()
            nn = Network(tuple(dd.wiring.slot_wires_list), tuple(dd.wiring.out_wires),
                         tuple(t.tensor_dim for t in dd.wiring.wire_types))
            tt = find_tree_elimination(nn, trials=2, seed=0)
            rec["decisions"].append({"inputs": spec, "ssa_path": [list(p) for p in tt.ssa_path],
                                     "sat": bool(reference_execute(dd, tt))})
        out["random"].append(rec)
        print("random", k, n, m, seed, len(rec["solutions"]), flush=True)
    # circuits: truth tables as relations over the diagram's ports
    for name, diag in [("half_adder", rb.half_adder), ("full_adder", rb.full_adder),
                       ("rc_adder(2)", rb.rc_adder(2)), ("wallace_multiplier(2)", rb.wallace_multiplier(2)),
                       ("wallace_multiplier(3)", rb.wallace_multiplier(3))]:
        f = diag.flatten()
        net = Network(tuple(f.wiring.slot_wires_list), tuple(f.wiring.out_wires),
                      tuple(t.tensor_dim for t in f.wiring.wire_types))
        tree = find_tree_elimination(net, trials=2, seed=0)
        res = reference_execute(f, tree)
        out["circuits"].append({"name": name, "ssa_path": [list(p) for p in tree.ssa_path],
                                "points": sorted("".join(map(str, p)) for p in res.to_set())})
        print("circuit", name, len(res.to_set()), flush=True)
    # SimpleContraction (reference executor, tensorsat/contractions/simple.py:251-292)
    for k, n, m, seed in [(3, 3, 6, 3), (3, 4, 4, 0)]:
        c = CNFInstance.random(k, n, m, rng=seed)
        d = c.diagram().flatten()
        L = len(d.blocks)
        path = [(0, 1)] * (L - 1)
        res = SimpleContraction(FinRel, d.wiring, path).contract(d)
        out["simple"].append({"k": k, "n": n, "m": m, "seed": seed, "linear_path": [list(p) for p in path],
                              "solutions": sorted("".join(map(str, p)) for p in res.to_set())})
    fa = rb.full_adder
    res = SimpleContraction(FinRel, fa.wiring, [(0, 1), (0, 1), (0, 1), (0, 1)]).contract(fa)
    out["simple"].append({"name": "full_adder", "linear_path": [[0, 1]] * 4,
                          "points": sorted("".join(map(str, p)) for p in res.to_set())})
    return out


def main() -> None:
    rng = np.random.default_rng(20250217)
    with open(os.path.join(HERE, "ops.json"), "w") as f:
        json.dump(gen_ops(rng), f, separators=(",", ":"))
    with open(os.path.join(HERE, "instances.json"), "w") as f:
        json.dump(gen_instances(), f, separators=(",", ":"))
    with open(os.path.join(HERE, "networks.json"), "w") as f:
        json.dump(gen_networks(), f, separators=(",", ":"))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
