"""BASELINE configs 2 / 4 / 5 at their benchmarked shapes (trees sliced to 2^30-entry intermediates), on sampled
SUB-CUBES whose exact answers are known and non-trivial (tensorsat_b200/planner/subcube.py):

* the pins follow a committed witness (a model of the CNF / the factors of the product), so the sub-cube contains a
  solution: SAT = 1, count >= 1;
* every answer — the witness sub-cube and its single-flip neighbours, which come out 0 or > 0 — is compared with an
  independent exact model counter (oracle/sharp_sat.py: DPLL + component caching, nothing shared with the tensor path);
* the same sub-cube through the byte representation (TSB_PLAN_NO_PACK), the generic kernel (TSB_PLAN_FORCE_GENERIC)
  and, for the multiplier, without the tensor-core lowering (TSB_PLAN_NO_GEMM) gives the same result.
"""

import numpy as np
import pytest

from oracle.sharp_sat import count_models, network_to_cnf
from tests.backend import bind_cuda

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, scope="module")
def _bind():
    bind_cuda()


def _run(f, sub, subleaves, mode, flags=0, slice_range=None):
    from tensorsat_b200 import _native

    net = sub.network
    plan = _native.Plan(net.inputs, net.output, net.size_dict, sub.ssa_path, sub.sliced_wires,
                        mode=_native.MODE_COUNT if mode == "count" else _native.MODE_SAT,
                        wire_is_free=f.wire_is_free, flags=flags)
    leaves = [_native.upload(np.ascontiguousarray(t, dtype=np.uint8)) for t in subleaves]
    out = _native.empty(plan.info["out_numel"], plan.mode)
    lo, hi = slice_range or (0, plan.info["n_slices"])
    for _ in range(2):   # the second execution replays the captured graph: same answer
        plan.execute(leaves, out, lo, hi)
    res = int(out.cpu().numpy().reshape(-1)[0])
    info, desc = plan.info, plan.describe()
    plan.close()
    return res, info, desc


def _cnf_of(f):
    from tensorsat_b200.lib.sat import CNFInstance

    n, m = (120, 511) if f.config == "c2" else (200, 852)
    return CNFInstance.random(3, n, m, rng=0)


def _neighbours(asg, k, seed):
    """The witness pins and k single-flip neighbours (seeded)."""
    rng = np.random.default_rng(seed)
    out = [dict(asg)]
    for w in rng.choice(sorted(asg), size=k, replace=False):
        a = dict(asg)
        a[int(w)] ^= 1
        out.append(a)
    return out


@pytest.mark.parametrize("config,keep,n_flips", [("c4", 4, 10), ("c2", 2, 5)])
def test_random_3sat_subcubes_match_the_exact_counter(config, keep, n_flips):
    from tensorsat_b200 import _native
    from tensorsat_b200.planner import subcube
    from tensorsat_b200.planner.sampling import fix_wires
    from tensorsat_b200.planner.tree import ContractionTree

    f = subcube.load(config)
    cnf = _cnf_of(f)
    assert all(any((f.extra["model"][abs(l) - 1] == 1) == (l > 0) for l in c) for c in cnf.clauses)   # the witness is a model
    sub, subleaves, asg = f.subcube(keep)
    assert all(w < f.n_vars for w in asg) and sub.num_slices == 2 ** keep
    seen = set()
    for a in _neighbours(asg, n_flips, seed=1):
        net, leaves = fix_wires(f.tree.network, f.leaves, a)
        t = ContractionTree(net, f.tree.ssa_path, sub.sliced_wires)
        want = count_models(f.n_vars, cnf.clauses, a)
        if config == "c4":   # boolean semiring: the benchmarked bit-packed plan
            got, info, _ = _run(f, t, leaves, "sat")
            assert info["n_packed_steps"] > 0 and info["log2_max_tensor"] == 30
            assert got == int(want > 0), (a, got, want)
        gotc, info, _ = _run(f, t, leaves, "count")
        assert gotc == want, (a, gotc, want)
        seen.add(want > 0)
    assert seen == {True, False}   # the neighbours are a real test: some sub-cubes are empty, some are not
    # the witness sub-cube through the other representations
    want = count_models(f.n_vars, cnf.clauses, asg)
    assert want >= 1
    mode = "sat" if config == "c4" else "count"
    ref = int(want > 0) if mode == "sat" else want
    assert _run(f, sub, subleaves, mode, flags=_native.PLAN_NO_PACK)[0] == ref
    assert _run(f, sub, subleaves, mode, flags=_native.PLAN_NO_FUSE)[0] == ref
    if config == "c4":
        assert _run(f, sub, subleaves, mode, flags=_native.PLAN_FORCE_GENERIC, slice_range=(0, 2))[0] == \
            _run(f, sub, subleaves, mode, slice_range=(0, 2))[0]
    # random pins (the round-1 sampling): the exact counter says what the executor must say
    rsub, rleaves, rasg = f.subcube(keep, witness=False, seed=0)
    assert _run(f, rsub, rleaves, mode)[0] == (count_models(f.n_vars, cnf.clauses, rasg) if mode == "count"
                                                else int(count_models(f.n_vars, cnf.clauses, rasg) > 0))


def test_multiplier_factoring_subcube():
    """16x16 Wallace multiplier with the product fixed to 65521 * 65519: the sub-cube pinned along the factors'
    wire values is satisfiable (answer 1), the same pins under a product that has no 16 x 16-bit factorisation
    (the prime 2^32 - 5) give 0, a single-flip neighbour is checked against the exact counter on the network's
    wire-level CNF, and the tensor-core path agrees with the CUDA-core path on one slice."""
    from tensorsat_b200 import _native
    from tensorsat_b200.planner import subcube
    from tensorsat_b200.planner.sampling import fix_wires
    from tensorsat_b200.planner.tree import ContractionTree

    f = subcube.load("c5")
    keep = 2
    sub, subleaves, asg = f.subcube(keep)
    got, info, desc = _run(f, sub, subleaves, "sat")
    assert got == 1
    assert any(o["kind"] == "gemm" for o in desc["ops"])          # tcgen05 steps are on this path
    pert = subcube.decode_leaves(f.extra["perturbed_leaves"])
    psub, pleaves, _ = f.subcube(keep, leaves=pert)
    assert _run(f, psub, pleaves, "sat")[0] == 0
    # wire-level CNF of the network: the counter's verdict for the witness pins and one neighbour
    wcnf = network_to_cnf(f.tree.network.inputs, f.leaves)
    assert count_models(f.tree.network.num_wires, wcnf, asg) == 1   # exactly the factorisation 65521 x 65519
    w = sorted(asg)[len(asg) // 2]
    a2 = dict(asg)
    a2[w] ^= 1
    net, leaves = fix_wires(f.tree.network, f.leaves, a2)
    assert _run(f, ContractionTree(net, f.tree.ssa_path, sub.sliced_wires), leaves, "sat")[0] == \
        int(count_models(f.tree.network.num_wires, wcnf, a2) > 0)
    # one slice: GEMM-lowered plan == plan restricted to the CUDA-core kernels
    a = _run(f, sub, subleaves, "sat", slice_range=(1, 2))[0]
    b = _run(f, sub, subleaves, "sat", flags=_native.PLAN_NO_GEMM | _native.PLAN_NO_PACK, slice_range=(1, 2))[0]
    assert a == b
