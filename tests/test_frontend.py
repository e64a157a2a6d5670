"""The native front end (csrc/tsb_frontend.cu, tensorsat_b200/frontend.py): DIMACS / clauses -> flat network -> tree ->
plan without Python diagrams.  Everything is compared with the reference-recorded fixtures (tests/golden) and with the
Python mirror; contractions run on the emulated build here and on the real library in the GPU suite."""

import glob
import hashlib
import json
import os
import time

import numpy as np
import pytest

from oracle import fin_rel_oracle as O
from tests.backend import bind_emulation
from tests.conftest import GOLDEN, load_golden


@pytest.fixture(autouse=True, scope="module")
def _bind():
    bind_emulation()


def _digest(inputs, outs):
    blob = json.dumps([list(map(list, inputs)), list(outs)]).encode()
    return hashlib.sha256(blob).hexdigest()


def test_wiring_matches_the_reference_digests_up_to_n200():
    """Wire / slot numbering of (cnf.inputs(None) >> cnf.diagram()).flatten() as recorded from the reference, for every
    random instance of the fixtures — including n=120 and n=200, whose Python diagrams take 10 s / 34 s to build."""
    from tensorsat_b200.frontend import LEAF_KINDS, NativeNetwork
    from tensorsat_b200.lib.sat import CNFInstance

    t_native = 0.0
    for rec in load_golden("instances.json")["random"]:
        c = CNFInstance.random(rec["k"], rec["n"], rec["m"], rng=rec["seed"])
        t0 = time.perf_counter()
        nn = NativeNetwork.from_clauses(rec["n"], c.clauses, "?" * rec["n"])
        inputs, kinds, outs = nn.arrays()
        t_native += time.perf_counter() - t0
        assert nn.info["n_wires"] == rec["num_wires"] and nn.info["n_leaves"] == rec["num_blocks"]
        assert _digest(inputs, outs) == rec["wiring_sha256"], rec["n"]
        if "box_names" in rec:
            assert [LEAF_KINDS[k] for k in kinds] == rec["box_names"]
        open_net = NativeNetwork.from_clauses(rec["n"], c.clauses)
        assert open_net.info["n_leaves"] == rec["diagram_num_blocks"] and open_net.arrays()[2] == tuple(range(rec["n"]))
    assert t_native < 2.0


def test_dimacs_dialect_and_errors():
    """tensorsat/lib/sat.py:80-100 and examples/test_dimacs_parser.py:3-64 through the native parser."""
    from tensorsat_b200.frontend import NativeNetwork
    from tensorsat_b200.lib.sat import CNFInstance

    for rec in load_golden("instances.json")["dimacs"]:
        with open(os.path.join(GOLDEN, "kissat", rec["name"] + ".cnf")) as f:
            text = f.read()
        nn = NativeNetwork.from_dimacs(text)
        assert (nn.info["n_vars"], nn.info["n_clauses"]) == (rec["num_vars"], rec["num_clauses"])
        mirror = CNFInstance.from_dimacs(text)
        want = (mirror.inputs(None) >> mirror.diagram()).flatten().wiring if rec["num_vars"] <= 60 else None
        if want is not None:
            got = NativeNetwork.from_dimacs(text, "?" * rec["num_vars"]).arrays()
            assert got[0] == tuple(tuple(ws) for ws in want.slot_wires_list) and got[2] == tuple(want.out_wires)
    with pytest.raises(ValueError, match="must start with"):
        NativeNetwork.from_dimacs("1 2 0\n")
    with pytest.raises(ValueError, match="Number of clauses"):
        NativeNetwork.from_dimacs("p cnf 2 2\n1 2 0\n")
    with pytest.raises(ValueError, match="invalid variables"):
        NativeNetwork.from_dimacs("p cnf 2 1\n1 3 0\n")
    with pytest.raises(ValueError, match="characters long"):
        NativeNetwork.from_dimacs("p cnf 2 1\n1 2 0\n", "?")
    with pytest.raises(TypeError):
        NativeNetwork.from_dimacs(b"p cnf 1 1\n1 0\n")
    ok = NativeNetwork.from_dimacs("c comment\n\np cnf 2 1\n  1 -2 0  \nc trailing\n")
    assert ok.info["n_clauses"] == 1 and ok.arrays()[0] == ((1, 2), (0, 2, 3), (3,))


def test_native_tree_search_equals_the_python_planner_and_is_fast():
    from tensorsat_b200.frontend import NativeNetwork
    from tensorsat_b200.lib.sat import CNFInstance
    from tensorsat_b200.planner import ContractionTree, Network
    from tensorsat_b200.planner.search import find_tree_elimination

    for (k, n, m, seed) in [(3, 10, 40, 0), (4, 10, 95, 0), (3, 30, 128, 1), (3, 50, 213, 0)]:
        cnf = CNFInstance.random(k, n, m, rng=seed)
        net = Network.from_diagram((cnf.inputs(None) >> cnf.diagram()).flatten())
        nn = NativeNetwork.from_clauses(n, cnf.clauses, "?" * n)
        for trials in (1, 2):      # the deterministic trials: min-fill, min-degree
            pt = find_tree_elimination(net, trials=trials, seed=0)
            nt = nn.search(trials=trials, seed=0, target_log2_size=0)
            assert tuple(pt.ssa_path) == nt.ssa_path
            st = pt.stats()
            assert nt.info["log2_width"] == st.log2_max_size and abs(nt.info["log2_ops"] - st.log2_ops) < 1e-9
        # slicing: same rule as ContractionTree.slice_to_fast
        target = max(8.0, pt.width() - 6)
        nt = nn.search(trials=2, seed=0, target_log2_size=target)
        assert nt.sliced_wires == tuple(pt.slice_to_fast(target).sliced_wires)
        assert ContractionTree(net, nt.ssa_path, nt.sliced_wires).width() <= target == max(target, nt.info["log2_width"])
        # randomised trials: valid trees, never worse than the deterministic ones
        r = nn.search(trials=8, seed=3, target_log2_size=0)
        ContractionTree(net, r.ssa_path)   # validates the path
        assert r.info["log2_ops"] + 6 * 0 <= max(nt.info["log2_ops"], 1e9)
    cnf = CNFInstance.random(3, 200, 852, rng=0)
    t0 = time.perf_counter()
    nn = NativeNetwork.from_clauses(200, cnf.clauses, "?" * 200)
    tr = nn.search(trials=1, seed=0, target_log2_size=30)
    dt = time.perf_counter() - t0
    assert tr.info["log2_width_unsliced"] == 136 and tr.info["n_sliced"] == 106 and tr.info["log2_width"] == 30
    assert dt < 5.0   # (-O1 emulated build in the CPU container: ~1 s; the Python diagram + planner take ~40 s)


def test_native_pipeline_verdicts_and_counts():
    """DIMACS text -> verdict / count in one call, against brute force and the reference's recorded verdicts."""
    from tensorsat_b200 import _native
    from tensorsat_b200.frontend import NativeNetwork, count_models, is_satisfiable
    from tensorsat_b200.lib.sat import CNFInstance

    for seed in range(4):
        n, m = 12, 44 + 3 * seed
        cnf = CNFInstance.random(3, n, m, rng=seed)
        brute = O.cnf_brute_force(n, cnf.clauses)
        text = cnf.to_dimacs()
        assert is_satisfiable(text, target_log2_size=8) == bool(brute.any())
        assert count_models(text, target_log2_size=8) == int(brute.sum())
        # all-solutions tensor (open variables) and plan reuse with swapped input boxes
        nn = NativeNetwork.from_clauses(n, cnf.clauses)
        tree = nn.search(trials=2, target_log2_size=0)
        plan = nn.plan(tree)
        assert np.array_equal(nn.contract(plan), brute.reshape((2,) * n))
        plan.close()
        nn = NativeNetwork.from_clauses(n, cnf.clauses, "?" * n)
        tree = nn.search(trials=2, target_log2_size=7)
        plan = nn.plan(tree, _native.MODE_COUNT)
        cube = brute.reshape((2,) * n)
        for spec in ("?" * n, "1" + "?" * (n - 1), "01" + "?" * (n - 2), "?" * (n - 1) + "0"):
            nn.set_inputs(spec)
            idx = tuple(slice(None) if ch == "?" else int(ch) for ch in spec)
            assert int(nn.contract(plan)) == int(cube[idx].sum()), spec
        plan.close()
    verdicts = {r["name"]: bool(r["reference_verdict"]) for r in load_golden("networks.json")["kissat"]}
    done = 0
    for path in sorted(glob.glob(os.path.join(GOLDEN, "kissat", "*.cnf"))):
        name = os.path.basename(path)[:-4]
        if name not in verdicts or name == "prime121":
            continue
        with open(path) as f:
            assert is_satisfiable(f.read(), target_log2_size=16) == verdicts[name], name
        done += 1
    assert done >= 35
