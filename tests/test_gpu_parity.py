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
