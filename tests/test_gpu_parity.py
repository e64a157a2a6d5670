"""GPU suite: the parity cases through the real libtsb.so on a B200, plus full-size properties."""

import numpy as np
import pytest

from tests import parity_cases
from tests.backend import bind_cuda
from tests.parity_cases import *  # noqa: F401,F403

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, scope="module")
def _bind():
    bind_cuda()
    old = dict(parity_cases.SCALE)
    parity_cases.SCALE["big_bits"] = 24
    yield
    parity_cases.SCALE.update(old)


def test_native_library_is_loaded():
    import ctypes

    from tensorsat_b200 import _native

    assert isinstance(_native.lib(), ctypes.CDLL) and _native._device == "cuda"
    assert _native.lib().tsb_version() == 100


def test_generic_kernel_vs_fast_kernels_large():
    """Device-side cross-check at a size the oracle cannot reach in seconds."""
    from tensorsat_b200.contractions.tree import TreeContraction
    from tensorsat_b200.lang.fin_rel import CudaCountRel, CudaFinRel
    from tensorsat_b200.lib.sat import CNFInstance
    from tensorsat_b200.planner import Network
    from tensorsat_b200.planner.search import find_tree_elimination

    cnf = CNFInstance.random(3, 30, 128, rng=1)
    d = (cnf.inputs(None) >> cnf.diagram()).flatten()
    tree = find_tree_elimination(Network.from_diagram(d), trials=4, seed=0)
    free = [w < 30 for w in d.wiring.wires]
    a = TreeContraction(CudaCountRel, d.wiring, tree=tree, wire_is_free=free).contract(d)
    b = TreeContraction(CudaCountRel, d.wiring, tree=tree, wire_is_free=free, force_generic=True).contract(d)
    s = TreeContraction(CudaCountRel, d.wiring, tree=tree.slice_to(14), wire_is_free=free).contract(d)
    assert int(a) == int(b) == int(s)
    sat = TreeContraction(CudaFinRel, d.wiring, tree=tree).contract(d)
    assert bool(sat) == (int(a) > 0)


def test_full_size_c1_properties():
    """BASELINE config 1 (3-SAT n=50 m=213) at full size: size-independent properties.

    * SAT verdict of the sliced and the unsliced plan agree;
    * #SAT: count(x0=0) + count(x0=1) == count, and count > 0 iff SAT;
    * if SAT, the model found by the reference's assignment search
      (notebooks/02-random-ksat.ipynb:707-722) satisfies every clause on the host.
    """
    from tensorsat_b200.contractions.tree import TreeContraction
    from tensorsat_b200.lang.fin_rel import CudaCountRel, CudaFinRel
    from tensorsat_b200.lib.sat import CNFInstance
    from tensorsat_b200.planner import Network
    from tensorsat_b200.planner.search import find_tree_elimination

    n = 50
    cnf = CNFInstance.random(3, n, 213, rng=0)
    wiring = cnf.contraction_wiring()
    tree = find_tree_elimination(Network.from_wiring(wiring), trials=6, seed=0).slice_to(28)
    free = [w < n for w in wiring.wires]
    sat_c = TreeContraction(CudaFinRel, wiring, tree=tree)
    cnt_c = TreeContraction(CudaCountRel, wiring, tree=tree, wire_is_free=free)

    def run(c, spec):
        return c.contract((cnf.inputs(spec) >> cnf.diagram()).flatten())

    is_sat = bool(run(sat_c, None))
    total = int(run(cnt_c, None))
    c0, c1 = int(run(cnt_c, "0" + "?" * (n - 1))), int(run(cnt_c, "1" + "?" * (n - 1)))
    assert c0 + c1 == total and (total > 0) == is_sat
    unsliced = TreeContraction(CudaFinRel, wiring, tree=tree.with_sliced(()))
    assert bool(run(unsliced, None)) == is_sat
    if is_sat:
        prefix = ""
        for _ in range(n):
            for bit in "01":
                if bool(run(sat_c, prefix + bit + "?" * (n - len(prefix) - 1))):
                    prefix += bit
                    break
            else:
                raise AssertionError("no extension of a satisfiable prefix is satisfiable")
        model = [int(ch) for ch in prefix]
        for clause in cnf.clauses:
            assert any((model[abs(l) - 1] == 1) == (l > 0) for l in clause)
        assert int(run(cnt_c, prefix)) == 1


def test_plan_executions_replay_as_cuda_graphs(monkeypatch):
    """From the second execution with the same slice range on, a plan replays its launch sequence as a CUDA graph
    (one host call): results stay exact when only the leaf VALUES change between replays, the per-kernel counters
    keep counting, TSB_PLAN_NO_GRAPH switches it off, and two slice ranges get two graphs."""
    from oracle import fin_rel_oracle as O
    from tensorsat_b200 import _native
    from tensorsat_b200.contractions.tree import TreeContraction
    from tensorsat_b200.lang.fin_rel import CudaCountRel, CudaFinRel
    from tensorsat_b200.lib.sat import CNFInstance
    from tensorsat_b200.planner import Network
    from tensorsat_b200.planner.search import find_tree_elimination

    monkeypatch.setenv("TSB_PACK_MIN_LOG2", "0")
    monkeypatch.setenv("TSB_FUSE_MIN_LOG2", "0")
    n = 18
    cnf = CNFInstance.random(3, n, 72, rng=3)
    brute = O.cnf_brute_force(n, cnf.clauses).reshape((2,) * n)
    wiring = cnf.contraction_wiring()
    tree = find_tree_elimination(Network.from_wiring(wiring), trials=2, seed=0)
    tree = tree.slice_to(max(tree.width() - 3, 6))
    assert tree.num_slices >= 4
    free = [w < n for w in wiring.wires]
    for cls, kw in ((CudaFinRel, {}), (CudaCountRel, {"wire_is_free": free})):
        c = TreeContraction(cls, wiring, tree=tree, batch_log2=10, **kw)
        specs = ["?" * n, "1" + "?" * (n - 1), "01" + "?" * (n - 2), "0" + "?" * (n - 1), "?" * n, "110" + "?" * (n - 3)]
        for spec in specs:
            flat = (cnf.inputs(spec) >> cnf.diagram()).flatten()
            sub = brute[tuple(int(ch) for ch in spec if ch != "?")]
            got = c.contract_from_host([b.tensor for b in flat.boxes])
            assert (int(got) == int(sub.sum())) if cls is CudaCountRel else (bool(got) == bool(sub.any())), spec
        d = c.plan().describe()
        assert d["graph_captures"] == 1 and d["graph_replays"] == len(specs) - 1
        # a second slice range: its own graph; halves combine to the whole
        flat = (cnf.inputs(None) >> cnf.diagram()).flatten()
        leaves = [b.tensor for b in flat.boxes]
        h = c.num_slices // 2
        for _ in range(3):
            lo = c.contract_from_host(leaves, (0, h))
            hi = c.contract_from_host(leaves, (h, c.num_slices))
            whole = (int(lo) + int(hi)) if cls is CudaCountRel else int(bool(lo) or bool(hi))
            assert whole == (int(brute.sum()) if cls is CudaCountRel else int(brute.any()))
        assert c.plan().describe()["graph_captures"] == 3
    # per-kernel counters count replayed launches too
    c = TreeContraction(CudaFinRel, wiring, tree=tree, batch_log2=10)
    flat = (cnf.inputs(None) >> cnf.diagram()).flatten()
    leaves = [b.tensor for b in flat.boxes]
    c.contract_from_host(leaves)
    _native.trace(True)
    before = _native.launch_count()
    c.contract_from_host(leaves)
    direct = (_native.launch_count() - before, {k: v for k, v in _native.kernel_counts().items() if v})   # capture + first replay
    _native.trace(True)
    before = _native.launch_count()
    c.contract_from_host(leaves)
    replay = (_native.launch_count() - before, {k: v for k, v in _native.kernel_counts().items() if v})
    _native.trace(False)
    assert c.plan().describe()["graph_replays"] == 2
    assert replay[0] == direct[0] > 0 and replay[1] == direct[1]
    assert replay[1]
    # switched off per plan
    import tensorsat_b200.contractions.tree as T

    c2 = TreeContraction(CudaFinRel, wiring, tree=tree, batch_log2=10)
    c2._flags |= _native.PLAN_NO_GRAPH
    for _ in range(3):
        assert bool(c2.contract_from_host(leaves)) == bool(brute.any())
    d2 = c2.plan().describe()
    assert d2["graph_captures"] == 0 and d2["graph_replays"] == 0
    assert T is not None


def test_plans_are_safe_across_streams():
    """Executions of one plan issued on different streams are serialised by the plan (event wait), and the arena
    hands blocks across streams in order: alternating streams gives the same answers."""
    import torch

    from oracle import fin_rel_oracle as O
    from tensorsat_b200.contractions.tree import TreeContraction
    from tensorsat_b200.lang.fin_rel import CudaCountRel
    from tensorsat_b200.lib.sat import CNFInstance
    from tensorsat_b200.planner import Network
    from tensorsat_b200.planner.search import find_tree_elimination

    n = 16
    cnf = CNFInstance.random(3, n, 60, rng=4)
    brute = O.cnf_brute_force(n, cnf.clauses).reshape((2,) * n)
    wiring = cnf.contraction_wiring()
    tree = find_tree_elimination(Network.from_wiring(wiring), trials=2, seed=0).slice_to(9)
    c = TreeContraction(CudaCountRel, wiring, tree=tree, wire_is_free=[w < n for w in wiring.wires])
    streams = [torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()]
    outs = []
    for i in range(9):
        spec = format(i % 4, "02b") + "?" * (n - 2)
        flat = (cnf.inputs(spec) >> cnf.diagram()).flatten()
        with torch.cuda.stream(streams[i % 3]):
            outs.append((spec, c.contract(flat)))
    torch.cuda.synchronize()
    for spec, box in outs:
        assert int(box) == int(brute[int(spec[0]), int(spec[1])].sum()), spec


def test_native_front_end_on_device():
    """DIMACS text -> verdict through csrc/tsb_frontend.cu on the GPU: the reference's regression set incl. prime121
    (n=329, 3817 leaves) and BASELINE config 1 at full size (n=50, native tree sliced to 2^28), against the recorded
    verdicts / the plan built by the Python mirror."""
    import glob
    import os

    from tensorsat_b200.frontend import is_satisfiable
    from tensorsat_b200.lib.sat import CNFInstance
    from tests.conftest import GOLDEN, load_golden

    verdicts = {r["name"]: bool(r["reference_verdict"]) for r in load_golden("networks.json")["kissat"]}
    for path in sorted(glob.glob(os.path.join(GOLDEN, "kissat", "*.cnf"))):
        name = os.path.basename(path)[:-4]
        if name in verdicts:
            with open(path) as f:
                assert is_satisfiable(f.read(), trials=2, target_log2_size=28) == verdicts[name], name
    cnf = CNFInstance.random(3, 50, 213, rng=0)
    assert is_satisfiable(cnf.to_dimacs(), trials=2, target_log2_size=28) is True   # (same verdict as bench.py's plan)


def test_assignment_search_at_n50_needs_a_fifth_of_the_contractions():
    """notebooks/02-random-ksat.ipynb:707-722 on BASELINE config 1: the lexicographically first model through
    AssignmentSearch (one plan, graph replays with swapped input leaves).  With unit propagation answering the
    implied decisions only 15 of the reference's 76 decisions cost a contraction; the model satisfies the formula and
    is the one the exact counter's self-reduction finds."""
    from oracle.sharp_sat import count_models
    from tensorsat_b200.contractions.tree import TreeContraction
    from tensorsat_b200.lang.fin_rel import CudaFinRel
    from tensorsat_b200.lib.sat import CNFInstance
    from tensorsat_b200.lib.search import AssignmentSearch
    from tensorsat_b200.planner import Network
    from tensorsat_b200.planner.search import find_tree_elimination

    n = 50
    cnf = CNFInstance.random(3, n, 213, rng=0)
    wiring = cnf.contraction_wiring()
    tree = find_tree_elimination(Network.from_wiring(wiring), trials=6, seed=0).slice_to(28)
    search = AssignmentSearch(cnf, TreeContraction(CudaFinRel, wiring, tree=tree))
    trace = []
    model = search.first_model(lambda spec, ok: trace.append((spec, ok)))
    assert model is not None and all(any((model[abs(l) - 1] == "1") == (l > 0) for l in c) for c in cnf.clauses)
    assert len(trace) == search.decisions + search.implied == 76 and search.decisions <= 19
    for spec, ok in trace[:: 7]:   # spot-check the trace against the exact counter
        assert ok == (count_models(n, cnf.clauses, {i: int(ch) for i, ch in enumerate(spec) if ch != "?"}) > 0)
    prefix = ""
    for _ in range(n):   # the counter's own self-reduction gives the same lexicographically first model
        prefix += "0" if count_models(n, cnf.clauses, {i: int(ch) for i, ch in enumerate(prefix + "0")}) > 0 else "1"
    assert prefix == model
