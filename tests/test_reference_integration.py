"""The device box bound into the REAL reference classes.

Runs in the build container only: the reference is imported from /root/reference through
oracle/ref_shims/load_reference.py (stand-ins for its absent pure-Python dependencies); the device side is the
same C-ABI code on the host emulation (tests/emu).  ``tensorsat_b200.integration.cuda_fin_rel.CudaFinRel`` derives
from the reference's own ``TensorLikeBox`` (created by its ``BoxMeta``), and the tests drive the reference's OWN
``SimpleContraction.contract``, ``_autoray.einsum / transpose / stack / shape``, ``Box.prod / from_wiring /
spider / scalar / union / stack`` and ``Box.__getitem__ / __add__ / __bool__`` over device boxes, comparing with
the reference's numpy ``FinRel`` on the same inputs.
"""

import os
import sys

import numpy as np
import pytest

from tests.backend import bind_emulation
from tests.conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shims"))
import load_reference  # noqa: E402

pytestmark = pytest.mark.skipif(not load_reference.available(), reason="the reference is only present in the build container")


@pytest.fixture(scope="module")
def ref():
    bind_emulation()
    load_reference.load()
    import tensorsat._autoray as ref_autoray
    from tensorsat.contractions.simple import SimpleContraction
    from tensorsat.diagrams import Box, Diagram, TensorLikeBox
    from tensorsat.lang.fin_rel import FinRel, FinSet
    from tensorsat.lib import bincirc
    from tensorsat.lib.sat import CNFInstance

    from tensorsat_b200.integration.cuda_fin_rel import CudaFinRel, to_device

    class NS:
        pass

    ns = NS()
    ns.__dict__.update(locals())
    return ns


def _rand(rng, r, p=0.5):
    return (rng.random((2,) * r) < p).astype(np.uint8)


def test_class_is_a_real_reference_box(ref):
    C = ref.CudaFinRel
    assert issubclass(C, ref.TensorLikeBox) and issubclass(C, ref.Box) and C.__final__
    assert type(C) is type(ref.FinRel)                       # created by the reference's BoxMeta
    assert C.can_be_contracted()
    assert ref.Box.class_join([C, C]) is C
    with pytest.raises(AssertionError):   # sibling classes: the reference's join (meta.py:440-465) finds none,
        ref.Box.class_join([C, ref.FinRel])   # so mixing host and device boxes in one einsum is refused by the reference itself
    import autoray                                            # the stand-in records BoxMeta's registration

    assert autoray._BACKENDS[C] == "tensorsat._autoray"
    with pytest.raises(TypeError):
        type("Sub", (C,), {})().shape  # final: no instantiable subclasses


def test_hooks_match_fin_rel(ref):
    rng = np.random.default_rng(0)
    C, F = ref.CudaFinRel, ref.FinRel
    for _ in range(20):
        nl, nr = int(rng.integers(1, 7)), int(rng.integers(1, 7))
        wires = list(range(8))
        lw = [int(w) for w in rng.choice(wires, nl, replace=False)]
        rw = [int(w) for w in rng.choice(wires, nr, replace=False)]
        allw = sorted(set(lw) | set(rw))
        ow = [int(w) for w in rng.permutation(allw)[: int(rng.integers(0, len(allw) + 1))]]
        a, b = _rand(rng, nl), _rand(rng, nr)
        want = F.contract2(F(a), lw, F(b), rw, ow)
        got = C.contract2(C(a), lw, C(b), rw, ow)
        assert got.shape == want.shape and np.array_equal(got.tensor, want.tensor)
        assert np.array_equal(C.contract2(C(a), lw, C(b), rw).tensor, F.contract2(F(a), lw, F(b), rw).tensor)
    t = _rand(rng, 5)
    for ports in ([4, 3, 2, 1, 0], [0, 2, 1, 3, 4], [0, 1, 2, 3], [1, 2]):
        assert np.array_equal(C(t).rewire(ports).tensor, F(t).rewire(ports).tensor)
    assert np.array_equal(C(t).rewire([0, 1, 0, 2, 3, 4]).tensor, F(t).rewire([0, 1, 0, 2, 3, 4]).tensor)
    u = _rand(rng, 5)
    assert np.array_equal((C(t) + C(u)).tensor, (F(t) + F(u)).tensor)
    assert np.array_equal(C(t)[1, :, 0].tensor, F(t)[1, :, 0].tensor)
    assert np.array_equal(C.stack([C(t), C(u), C(t)], 2).tensor, F.stack([F(t), F(u), F(t)], 2).tensor)
    assert np.array_equal(C.union([C(t), C(u)]).tensor, F.union([F(t), F(u)]).tensor)
    assert np.array_equal(C.prod([C(t[0, 0]), C(u[1, 1])]).tensor, F.prod([F(t[0, 0]), F(u[1, 1])]).tensor)
    for n in (1, 2, 3):
        assert np.array_equal(C.spider(ref.FinSet(3), n).tensor, F.spider(ref.FinSet(3), n).tensor)
    assert bool(C.scalar(True)) is True and bool(C.scalar(False)) is False
    with pytest.raises(ValueError):
        bool(C(t))
    with pytest.raises(NotImplementedError):
        C(t)[slice(0, 1), 0]
    with pytest.raises(ValueError):
        C(np.full((2,), 2, np.uint8))
    assert C(t) == C(t.copy()) and hash(C(t)) == hash(C(t.copy())) and C(t) != C(u)
    assert C(t).to_fin_rel() == F(t) and C.from_fin_rel(F(t)) == C(t)


def test_reference_autoray_module_drives_device_boxes(ref):
    """tensorsat/_autoray.py:31-80 unchanged, over device boxes."""
    rng = np.random.default_rng(1)
    C, F, ab = ref.CudaFinRel, ref.FinRel, ref.ref_autoray
    a, b = _rand(rng, 4), _rand(rng, 3)
    for eq in ("abcd,cde->abe", "abcd,dca->b", "abcd,efg->gabf", "abcd,bcd->"):
        assert np.array_equal(ab.einsum(eq, C(a), C(b)).tensor, ab.einsum(eq, F(a), F(b)).tensor), eq
    assert np.array_equal(ab.transpose(C(a), (3, 1, 0, 2)).tensor, ab.transpose(F(a), (3, 1, 0, 2)).tensor)
    assert ab.shape(C(a)) == ab.shape(F(a)) == (2, 2, 2, 2)
    assert np.array_equal(ab.stack([C(b), C(b)], 1).tensor, ab.stack([F(b), F(b)], 1).tensor)
    with pytest.raises(NotImplementedError):
        ab.einsum("ab->a", C(a[0, 0]))
    # a whole tree stepped through the reference's einsum, as cotengra's traversal does (cotengra.py:120-126)
    cnf = ref.CNFInstance.random(3, 8, 30, rng=2)
    d = cnf.diagram().flatten()
    dd = ref.to_device(d)
    from tensorsat_b200.planner import Network
    from tensorsat_b200.planner.search import find_tree_elimination

    net = Network([tuple(ws) for ws in d.wiring.slot_wires_list], tuple(d.wiring.out_wires),
                  tuple(t.tensor_dim for t in d.wiring.wire_types))
    tree = find_tree_elimination(net, trials=1, seed=0)
    letters = [chr(ord("a") + i) if i < 26 else chr(ord("A") + i - 26) if i < 52 else chr(192 + i) for i in range(net.num_wires)]
    for boxes in (list(d.boxes), list(dd.boxes)):
        pool = {i: (bx, net.inputs[i]) for i, bx in enumerate(boxes)}
        for st in tree.steps:
            (lb, lw), (rb, rw) = pool.pop(st.lhs), pool.pop(st.rhs)
            ow = tuple(sorted(st.out_wires))
            eq = "".join(letters[w] for w in lw) + "," + "".join(letters[w] for w in rw) + "->" + "".join(letters[w] for w in ow)
            pool[st.out] = (ab.einsum(eq, lb, rb), ow)
        (res, ow), = pool.values()
        perm = [ow.index(w) for w in net.output]
        res = ab.transpose(res, perm)
        if boxes[0].__class__ is F:
            want = res.tensor
        else:
            assert isinstance(res, C) and np.array_equal(res.tensor, want)


def test_reference_simple_contraction_on_device_boxes(ref):
    """tensorsat/contractions/simple.py:251-292 unchanged: box_class = the device class."""
    C, F, S = ref.CudaFinRel, ref.FinRel, ref.SimpleContraction
    for k, n, m, seed in [(3, 3, 3, 1), (3, 4, 3, 2), (2, 3, 3, 1), (3, 4, 4, 0)]:   # (keeps every wire to the end: tiny)
        d = ref.CNFInstance.random(k, n, m, rng=seed).diagram().flatten()
        path = [(0, 1)] * (len(d.blocks) - 1)
        want = S(F, d.wiring, path).contract(d)
        got = S(C, d.wiring, path).contract(ref.to_device(d))
        assert isinstance(got, C) and got.shape == want.shape and np.array_equal(got.tensor, want.tensor)
    fa = ref.bincirc.full_adder
    want = S(F, fa.wiring, [(0, 1)] * 4).contract(fa)
    got = S(C, fa.wiring, [(0, 1)] * 4).contract(ref.to_device(fa))
    assert got.to_set() == want.to_set() and len(got.to_set()) == 8
    with pytest.raises(ValueError):   # host boxes in a device contraction: the reference's own validation
        S(C, fa.wiring, [(0, 1)] * 4).contract(fa)
    # Box.from_wiring (diagrams.py:750-775) through spider / prod / _rewire on the device
    w = ref.CNFInstance.random(3, 3, 2, rng=0).diagram().flatten().wiring
    assert np.array_equal(C.from_wiring(w).tensor, F.from_wiring(w).tensor)
