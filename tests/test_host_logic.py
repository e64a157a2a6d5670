"""Host-side logic: instance generators, DIMACS parser, diagram mirror, planner, slice partition."""

import hashlib
import json
import os

import numpy as np
import pytest

from tensorsat_b200.lib import bincirc as bc
from tensorsat_b200.lib.sat import CNFInstance
from tensorsat_b200.planner import ContractionTree, Network, find_tree, linear_to_ssa, ssa_to_linear
from tensorsat_b200.planner.search import find_tree_elimination
from tests.conftest import GOLDEN, load_golden


def _digest(wiring):
    blob = json.dumps([list(map(list, wiring.slot_wires_list)), list(wiring.out_wires)]).encode()
    return hashlib.sha256(blob).hexdigest()


def test_random_instances_match_reference_generator():
    """CNFInstance.random RNG stream + bintree wiring (tensorsat/lib/sat.py:43-78,248-262)."""
    assert CNFInstance.random(3, 4, 4, rng=0).clauses == ((2, 3, 4), (-4, -3, -1), (-3, -2, 4), (-4, -3, 1))  # G1
    assert CNFInstance.random(3, 3, 6, rng=3).clauses == (
        (-3, -1, 2), (-3, 1, 2), (1, 2, 3), (-2, -1, 3), (-1, 2, 3), (-2, 1, 3))  # G2
    for rec in load_golden("instances.json")["random"]:
        c = CNFInstance.random(rec["k"], rec["n"], rec["m"], rng=rec["seed"])
        assert hashlib.sha256(json.dumps(c.clauses).encode()).hexdigest() == rec["clauses_sha256"]
        if rec["n"] > 50:
            continue
        d = (c.inputs(None) >> c.diagram()).flatten()
        assert d.wiring.num_wires == rec["num_wires"] and len(d.blocks) == rec["num_blocks"]
        assert _digest(d.wiring) == rec["wiring_sha256"]
        assert len(c.diagram().blocks) == rec["diagram_num_blocks"]
        if rec["n"] <= 20:
            assert _digest(c.contraction_wiring()) == rec["contraction_wiring_sha256"]
        if "slot_wires_list" in rec:
            assert [list(ws) for ws in d.wiring.slot_wires_list] == rec["slot_wires_list"]
            assert [b.name for b in d.boxes] == rec["box_names"]


def test_circuits_match_reference_wirings():
    made = {"half_adder": bc.half_adder, "full_adder": bc.full_adder}
    g = load_golden("instances.json")
    for rec in g["circuits"]:
        name = rec["name"]
        if name not in made:
            fn, arg = name[:-1].split("(")
            made[name] = getattr(bc, fn)(int(arg))
        f = made[name].flatten()
        assert (f.wiring.num_wires, len(f.blocks), f.num_ports) == (rec["num_wires"], rec["num_blocks"], rec["num_ports"])
        assert _digest(f.wiring) == rec["wiring_sha256"] and list(f.wiring.out_wires) == rec["out_wires"]
    for name, tensor in g["gates"].items():
        assert getattr(bc, name).tensor.tolist() == tensor  # G6 and friends


def test_dimacs_parser_golden():
    """examples/test_dimacs_parser.py:3-64 plus every file of the SAT test set."""
    for rec in load_golden("instances.json")["dimacs"]:
        with open(os.path.join(GOLDEN, "kissat", rec["name"] + ".cnf")) as f:
            c = CNFInstance.from_dimacs(f.read())
        assert (c.num_vars, len(c.clauses)) == (rec["num_vars"], rec["num_clauses"])
        assert hashlib.sha256(json.dumps(c.clauses).encode()).hexdigest() == rec["clauses_sha256"]
        if "clauses" in rec:
            assert [list(cl) for cl in c.clauses] == rec["clauses"]
        assert CNFInstance.from_dimacs(c.to_dimacs()).clauses == c.clauses
    with pytest.raises(ValueError):
        CNFInstance.from_dimacs("c nothing\n1 2 0\n")
    with pytest.raises(ValueError):
        CNFInstance.from_dimacs("p cnf 2 2\n1 2 0\n")
    with pytest.raises(ValueError):
        CNFInstance.from_dimacs("p cnf 1 1\n1 2 0\n")


def test_inputs_and_validation():
    c = CNFInstance.random(3, 5, 8, rng=1)
    d = c.inputs("01?0?")
    assert [b.name for b in d.boxes] == ["bit_0", "bit_1", "bit_unk", "bit_0", "bit_unk"]
    assert c.inputs(5).wiring == c.inputs("00101").wiring
    with pytest.raises(ValueError):
        c.inputs("01")
    with pytest.raises(ValueError):
        c.inputs("01x0?")
    with pytest.raises(ValueError):
        CNFInstance.random(3, 2, 4)
    with pytest.raises(ValueError):
        CNFInstance(2, [[1, 3]])


def test_path_conventions_roundtrip():
    rng = np.random.default_rng(0)
    for L in (2, 3, 7, 20):
        ids = list(range(L))
        ssa = []
        nxt = L
        while len(ids) > 1:
            a, b = rng.choice(len(ids), size=2, replace=False)
            x, y = ids[a], ids[b]
            ssa.append((x, y))
            ids = [i for i in ids if i not in (x, y)] + [nxt]
            nxt += 1
        lin = ssa_to_linear(ssa, L)
        assert linear_to_ssa(lin, L) == tuple(ssa)


def test_tree_stats_match_reference_cost_model():
    """cotengra statistics the reference prints for 3-SAT n=4 m=4 rng=0 (02-random-ksat.ipynb:310-314):
    N = 19 tensors, max tensor 16 B.  (ops / write depend on the tree: 232 ops, 132 B there.)"""
    c = CNFInstance.random(3, 4, 4, rng=0)
    d = c.diagram().flatten()
    tree = find_tree_elimination(Network.from_diagram(d), trials=4, seed=0)
    st = tree.stats()
    assert len(d.blocks) == 19 and st.num_contractions == 18
    assert st.log2_max_size == 4.0
    assert 2.0 ** st.log2_write <= 2 * 132 and 2.0 ** st.log2_ops <= 4 * 232
    t2 = ContractionTree.loads(tree.dumps())
    assert t2.ssa_path == tree.ssa_path and t2.stats().log2_ops == st.log2_ops


def test_tree_search_and_slicing():
    c = CNFInstance.random(3, 30, 128, rng=0)
    d = (c.inputs(None) >> c.diagram()).flatten()
    net = Network.from_diagram(d)
    e = find_tree_elimination(net, trials=4, seed=0)
    g = find_tree(net, trials=2, seed=0)
    assert e.stats().num_contractions == g.stats().num_contractions == net.num_leaves - 1
    assert e.width() <= 24
    s = e.slice_to(e.width() - 4)
    assert s.num_slices >= 16 and s.width() <= e.width() - 4
    assert s.stats().log2_ops < e.stats().log2_ops + 2.0
    with pytest.raises(ValueError):
        ContractionTree(net, e.ssa_path[:-1])
    with pytest.raises(ValueError):
        ContractionTree(net, e.ssa_path, sliced_wires=(1, 1))
    with pytest.raises(ValueError):
        ContractionTree(net, ((0, 0),) + e.ssa_path[1:])


def test_slice_to_on_non_binary_extents():
    """slice_to compares log2 SIZES (not wire counts) when it picks the largest intermediates: a ring of four
    extent-4 wires is sliceable to 2^2-entry intermediates."""
    net = Network(((0, 1), (1, 2), (2, 3), (3, 0)), (), (4, 4, 4, 4))
    tree = ContractionTree(net, [(0, 1), (2, 3), (4, 5)])
    assert tree.width() == 4.0
    sliced = tree.slice_to(2)
    assert sliced.width() <= 2.0 and sliced.num_slices == 16
    mixed = ContractionTree(Network(((0, 1), (1, 2), (2, 0)), (), (3, 2, 6)), [(0, 1), (3, 2)])
    assert mixed.slice_to(2).width() <= 2.0


def test_annealed_elimination_order_is_valid_and_never_worse():
    """The opt-in annealing of the order's tail (planner/search.py) is deterministic for a seed, returns a valid
    tree over the same network, and is kept only when the real objective improves."""
    from tensorsat_b200.planner.search import _objective

    c = CNFInstance.random(3, 30, 128, rng=1)
    net = Network.from_diagram((c.inputs(None) >> c.diagram()).flatten())
    base = find_tree_elimination(net, trials=2, seed=0)
    a = find_tree_elimination(net, trials=2, seed=0, anneal_iters=3000, anneal_tail=30, anneal_restarts=2)
    b = find_tree_elimination(net, trials=2, seed=0, anneal_iters=3000, anneal_tail=30, anneal_restarts=2)
    assert a.ssa_path == b.ssa_path
    assert a.network == net and a.stats().num_contractions == net.num_leaves - 1
    assert _objective(a, "combo", None) <= _objective(base, "combo", None)


def test_slice_blocks_partition_all_slices():
    from tensorsat_b200.distributed import slice_block

    for n in (1, 5, 64, 100):
        for w in (1, 2, 3, 8):
            blocks = [slice_block(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in blocks) - min(b - a for a, b in blocks) <= 1
    with pytest.raises(ValueError):
        slice_block(4, 2, 2)


def test_subcube_plan_files_are_consistent():
    """bench_plans/subcube_*.json (tools/make_subcube_plans.py): the committed witnesses are solutions, the pins of a
    sub-cube follow them, the sub-cubes are non-empty by the independent exact counter, and the trees fit the
    2^30-entry budget the benchmarks quote."""
    from oracle.sharp_sat import count_models
    from tensorsat_b200.planner import subcube

    for config, n, m in (("c2", 120, 511), ("c4", 200, 852)):
        f = subcube.load(config)
        cnf = CNFInstance.random(3, n, m, rng=0)
        model = f.extra["model"]
        assert all(any((model[abs(l) - 1] == 1) == (l > 0) for l in c) for c in cnf.clauses)
        assert f.tree.network.num_leaves == len(f.leaves) and f.tree.width() <= 30
        sub, subleaves, asg = f.subcube(4)
        assert sub.num_slices == 16 and all(model[w] == v for w, v in asg.items())
        assert count_models(n, cnf.clauses, asg) >= 1
        assert len(asg) == len(f.tree.sliced_wires) - 4 and sub.width() <= 30
    assert count_models(120, CNFInstance.random(3, 120, 511, rng=0).clauses) == 195572   # #SAT of BASELINE config 2
    f = subcube.load("c5")
    assert f.extra["factors"] == [65521, 65519] and f.tree.width() <= 30
    sub, subleaves, asg = f.subcube(2)
    for ws, t in zip(sub.network.inputs, subleaves):
        assert t.ndim == len(ws)


def test_bench_script_is_whole():
    """bench.py refers only to functions it defines (a cheap guard: the driver runs it unattended)."""
    import ast

    src = open(os.path.join(os.path.dirname(GOLDEN), "..", "bench.py")).read()
    tree = ast.parse(src)
    defined = {n.name for n in tree.body if isinstance(n, ast.FunctionDef)} | {n.name for n in tree.body if isinstance(n, ast.ClassDef)}
    called = {n.func.id for n in ast.walk(tree) if isinstance(n, ast.Call) and isinstance(n.func, ast.Name)}
    import builtins

    local_names = {n.id for n in ast.walk(tree) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store)}
    local_names |= {a.arg for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) for a in n.args.args}
    local_names |= {n.name for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)}
    local_names |= {(a.asname or a.name).split(".")[0] for n in ast.walk(tree) if isinstance(n, (ast.Import, ast.ImportFrom)) for a in n.names}
    missing = {c for c in called if c not in defined and c not in local_names and not hasattr(builtins, c)}
    assert not missing, missing
    for fn in ("load_tree", "cpu_sample", "run_reference", "micro_rooflines", "c4_subcube_leg", "main"):
        assert fn in defined


def test_host_helper_packs_leaves_like_join():
    """csrc/tsb_pyhost.c (the CPython helper behind contract_from_host): one pass over the leaves equals the joined
    C-order bytes; anything it cannot take (non-contiguous, not a buffer, too long) is refused with -1."""
    import numpy as np

    from tensorsat_b200 import build

    if build.build_host_helper() is None:
        pytest.skip("no C compiler / Python headers: the pure-Python packing path is used")
    from tensorsat_b200 import _tsb_host

    rng = np.random.default_rng(5)
    leaves = [rng.integers(0, 2, size=(2,) * int(r), dtype=np.uint8) for r in rng.integers(0, 4, size=300)]
    want = b"".join(a.tobytes() for a in leaves)
    dst = np.zeros(len(want), dtype=np.uint8)
    assert _tsb_host.pack_into(leaves, dst) == len(want) and dst.tobytes() == want
    assert _tsb_host.pack_into([a.tobytes() for a in leaves], dst) == len(want) and dst.tobytes() == want
    assert _tsb_host.pack_into(leaves + [np.zeros((4, 4), np.uint8)[:, ::2]], np.zeros(len(want) + 64, np.uint8)) == -1
    assert _tsb_host.pack_into(leaves + [3], np.zeros(len(want) + 64, np.uint8)) == -1
    assert _tsb_host.pack_into(leaves + [np.zeros(8, np.uint8)], dst) == -1


def test_two_slice_lanes_on_the_emulated_build():
    """TSB_SLICE_LANES=2 (slices alternate between two streams, the second with its own copy of the per-slice nodes):
    the knobs are read once per process, so the sliced packed-plan parity cases run in a child process with them set
    (TSB_WAVE_STREAMS > 0 lets direct launches fork as a captured graph would)."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TSB_SLICE_LANES="2", TSB_WAVE_STREAMS="2")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_emu_parity.py", "-q", "-x", "-p", "no:cacheprovider",
                        "-k", "partial_outputs and (24-96-4-12-7-2-3 or 22-90-6-13-6-2-1)"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
