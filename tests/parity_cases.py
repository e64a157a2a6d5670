"""
Parity cases shared by the CPU suite (C-ABI code running on the host emulation, tests/emu) and
the GPU suite (the real libtsb.so on a B200).  Every case calls the product through its public
box / contraction API (hence through the C-ABI) and compares bit-for-bit with the oracle.
"""

from __future__ import annotations

import numpy as np
import pytest

from oracle import fin_rel_oracle as O
from tensorsat_b200.contractions.simple import SimpleContraction
from tensorsat_b200.contractions.tree import TreeContraction
from tensorsat_b200.diagrams import FinSet
from tensorsat_b200.lang.fin_rel import CudaCountRel, CudaFinRel
from tensorsat_b200.lib import bincirc as bc
from tensorsat_b200.lib.sat import CNFInstance
from tensorsat_b200.planner import ContractionTree, Network
from tensorsat_b200.planner.search import find_tree_elimination

from tests.conftest import load_golden

SCALE = {"big_bits": 15}  # raised by the GPU suite


def _arr(x, shape=None, dtype=np.uint8):
    a = np.array(x, dtype=dtype)
    return a.reshape(shape) if shape is not None else a


def _rand01(rng, shape, p=0.5):
    return (rng.random(shape) < p).astype(np.uint8)


# ------------------------------------------------------------------------------------ golden ops
def test_contract2_golden():
    """FinRel.contract2 outputs recorded from the reference (where it is well defined)."""
    n = 0
    for c in load_golden("ops.json"):
        if c["op"] != "contract2":
            continue
        lhs, rhs = _arr(c["lhs"], c["lhs_shape"]), _arr(c["rhs"], c["rhs_shape"])
        ow = c["out_wires"]
        got = CudaFinRel.contract2(CudaFinRel(lhs), c["lhs_wires"], CudaFinRel(rhs), c["rhs_wires"], ow)
        ow_eff = ow if ow is not None else sorted(set(c["lhs_wires"]).symmetric_difference(c["rhs_wires"]))
        want = O.contract2(lhs, c["lhs_wires"], rhs, c["rhs_wires"], ow_eff)
        assert got.tensor.dtype == np.uint8
        assert got.tensor.shape == want.shape
        assert np.array_equal(got.tensor, want), c
        n += 1
        # counting semiring on the same case
        gotc = CudaCountRel.contract2(CudaCountRel(lhs), c["lhs_wires"], CudaCountRel(rhs), c["rhs_wires"], ow)
        wantc = O.contract2(lhs, c["lhs_wires"], rhs, c["rhs_wires"], ow_eff, mode="count")
        assert gotc.tensor.dtype == np.int64 and np.array_equal(gotc.tensor, wantc)
    assert n >= 50


def test_rewire_golden():
    for c in load_golden("ops.json"):
        if c["op"] != "rewire":
            continue
        t = _arr(c["tensor"], c["shape"])
        got = CudaFinRel(t).rewire(c["ports"])
        want = O.rewire(t, c["ports"])
        assert got.tensor.shape == want.shape and np.array_equal(got.tensor, want), c
        gotc = CudaCountRel(t).rewire(c["ports"])
        assert np.array_equal(gotc.tensor, O.rewire(t, c["ports"], mode="count"))


def test_small_ops_golden():
    for c in load_golden("ops.json"):
        if c["op"] == "spider":
            got = CudaFinRel.spider(FinSet(c["size"]), c["num_ports"])
            assert np.array_equal(got.tensor, _arr(c["res"]))
        elif c["op"] == "scalar":
            got = CudaFinRel.scalar(c["value"])
            assert got.shape == () and bool(got) == c["bool"]
            assert np.array_equal(got.tensor, _arr(c["res"]))
        elif c["op"] == "slice_union_stack":
            a, b = CudaFinRel(_arr(c["a"], c["shape"])), CudaFinRel(_arr(c["b"], c["shape"]))
            idx = tuple(slice(None) if i is None else i for i in c["idx"])
            assert np.array_equal(a[idx].tensor, _arr(c["getitem"], c["getitem_shape"]))
            assert np.array_equal((a + b).tensor, _arr(c["union"]))
            assert np.array_equal(CudaFinRel.stack([a, b, a], c["port"]).tensor, _arr(c["stack"]))
        elif c["op"] == "einsum":
            from tensorsat_b200 import autoray_backend as ab

            got = ab.einsum(c["eq"], CudaFinRel(_arr(c["lhs"])), CudaFinRel(_arr(c["rhs"])))
            assert np.array_equal(got.tensor, _arr(c["res"]))
        elif c["op"] == "transpose":
            from tensorsat_b200 import autoray_backend as ab

            got = ab.transpose(CudaFinRel(_arr(c["tensor"])), c["perm"])
            assert np.array_equal(got.tensor, _arr(c["res"]))


def test_mod256_quirk_documented():
    """The product implements count > 0; the reference's np.sign(dtype=uint8) says 0 for 256 witnesses."""
    ones8 = CudaFinRel(np.ones((2,) * 8, dtype=np.uint8))
    w = list(range(8))
    assert bool(CudaFinRel.contract2(ones8, w, ones8, w, [])) is True
    assert int(CudaCountRel.contract2(CudaCountRel.from_rel(ones8), w, CudaCountRel.from_rel(ones8), w, [])) == 256
    q = [c for c in load_golden("ops.json") if c["op"] == "contract2_quirk" and c["k"] == 8][0]
    assert q["res"] == 0
    assert int(O.contract2(np.ones((2,) * 8, np.uint8), w, np.ones((2,) * 8, np.uint8), w, [], ref_quirk_mod256=True)) == 0


# ------------------------------------------------------------------------------------ random contract2
def _random_contract_case(rng, nl, nr, n_shared, n_out_drop, dims=(2,)):
    n_wires = nl + nr - n_shared
    wires = list(rng.permutation(n_wires))
    shared = wires[:n_shared]
    lw = shared + wires[n_shared : n_shared + (nl - n_shared)]
    rw = shared + wires[nl:]
    lw = [int(w) for w in rng.permutation(lw)]
    rw = [int(w) for w in rng.permutation(rw)]
    allw = sorted(set(lw) | set(rw))
    drop = set(int(w) for w in rng.choice(allw, size=min(n_out_drop, len(allw)), replace=False)) if n_out_drop else set()
    ow = [int(w) for w in rng.permutation([w for w in allw if w not in drop])]
    wdim = {w: int(rng.choice(dims)) for w in allw}
    return lw, rw, ow, wdim


@pytest.mark.parametrize("seed", range(12))
def test_contract2_random_medium(seed):
    """Ranks large enough for the merge / permute kernels, arbitrary axis orders."""
    rng = np.random.default_rng(1000 + seed)
    nl, nr = int(rng.integers(6, 13)), int(rng.integers(6, 13))
    ns = int(rng.integers(1, min(nl, nr)))
    lw, rw, ow, wdim = _random_contract_case(rng, nl, nr, ns, int(rng.integers(0, 5)))
    lhs = _rand01(rng, [wdim[w] for w in lw], 0.3)
    rhs = _rand01(rng, [wdim[w] for w in rw], 0.3)
    got = CudaFinRel.contract2(CudaFinRel(lhs), lw, CudaFinRel(rhs), rw, ow)
    assert np.array_equal(got.tensor, O.contract2(lhs, lw, rhs, rw, ow))
    gotc = CudaCountRel.contract2(CudaCountRel(lhs), lw, CudaCountRel(rhs), rw, ow)
    assert np.array_equal(gotc.tensor, O.contract2(lhs, lw, rhs, rw, ow, mode="count"))


@pytest.mark.parametrize("seed", range(6))
def test_contract2_random_mixed_dims(seed):
    rng = np.random.default_rng(2000 + seed)
    lw, rw, ow, wdim = _random_contract_case(rng, 5, 5, 2, 2, dims=(2, 3, 4))
    lhs = _rand01(rng, [wdim[w] for w in lw])
    rhs = _rand01(rng, [wdim[w] for w in rw])
    got = CudaFinRel.contract2(CudaFinRel(lhs), lw, CudaFinRel(rhs), rw, ow)
    assert np.array_equal(got.tensor, O.contract2(lhs, lw, rhs, rw, ow))


def test_contract2_ordered_big():
    """An 'absorb' step in plan order (no permutes): big x small with batch wires, k = 1."""
    rng = np.random.default_rng(7)
    nb = SCALE["big_bits"]
    big_w = list(range(nb + 1))          # wire 0 is summed, the rest kept
    small_w = [0, 3, 5, nb + 1, nb + 2]  # shares 0 (summed), 3, 5 (batch), adds two wires
    ow = sorted(set(big_w[1:]) | set(small_w[1:]))
    big = _rand01(rng, (2,) * len(big_w), 0.4)
    small = _rand01(rng, (2,) * len(small_w), 0.6)
    got = CudaFinRel.contract2(CudaFinRel(big), big_w, CudaFinRel(small), small_w, ow)
    want = O.contract2(big, big_w, small, small_w, ow)
    assert np.array_equal(got.tensor, want)
    gotc = CudaCountRel.contract2(CudaCountRel(big), big_w, CudaCountRel(small), small_w, ow)
    assert np.array_equal(gotc.tensor, O.contract2(big, big_w, small, small_w, ow, mode="count"))


@pytest.mark.parametrize("tl,tr", [("u8", "u8"), ("u8", "i32"), ("i32", "u8"), ("i32", "i32")])
def test_counting_merge_int32_kernels(tl, tr):
    """The compile-time specialised counting kernels (int32 results) through tsb_contract2, all operand type pairs."""
    import torch

    from tensorsat_b200 import _native

    rng = np.random.default_rng(11)
    dt = {"u8": (np.uint8, torch.uint8, 4), "i32": (np.int32, torch.int32, 3000)}
    for trial in range(6):
        R = int(rng.integers(12, SCALE["big_bits"] - 1))
        out_w = list(range(2, R + 2))                       # wires 0, 1 are summed; all wires in one global order
        lw = [0, 1] + sorted(int(w) for w in rng.choice(out_w, size=int(rng.integers(2, R)), replace=False))
        rest = [w for w in out_w if w not in lw]
        rw = [0] + ([1] if trial % 2 else []) + sorted(set(rest) | set(int(w) for w in rng.choice(out_w, size=int(rng.integers(1, R)), replace=False)))
        L = rng.integers(0, dt[tl][2], size=(2,) * len(lw)).astype(dt[tl][0])
        Rr = rng.integers(0, dt[tr][2], size=(2,) * len(rw)).astype(dt[tr][0])
        dev = _native.device()
        Ld, Rd = torch.from_numpy(L.reshape(-1)).to(dev), torch.from_numpy(Rr.reshape(-1)).to(dev)
        Od = torch.empty(1 << R, dtype=torch.int32, device=dev)
        _native.contract2(Ld, [2] * len(lw), lw, Rd, [2] * len(rw), rw, Od, [2] * R, out_w, 1)
        want = O.einsum2(L.astype(np.int64), lw, Rr.astype(np.int64), rw, out_w)
        assert want.max() < 2**31
        assert np.array_equal(Od.cpu().numpy().reshape((2,) * R).astype(np.int64), want)


@pytest.mark.parametrize("tl", ["u8", "i32", "i64"])
@pytest.mark.parametrize("tr", ["u8", "i32", "i64"])
def test_counting_merge_int64_kernel(tl, tr):
    """The int64-result counting kernel (two outputs per thread) through tsb_contract2, all nine operand type pairs,
    with 0..3 summed wires and random lane masks."""
    import torch

    from tensorsat_b200 import _native

    rng = np.random.default_rng(12)
    dt = {"u8": (np.uint8, 4), "i32": (np.int32, 1 << 20), "i64": (np.int64, 1 << 38)}
    for trial in range(5):
        R = int(rng.integers(6, SCALE["big_bits"] - 2))
        nk = trial % 4
        out_w = list(range(nk, R + nk))
        ks = list(range(nk))
        lw = ks + sorted(int(w) for w in rng.choice(out_w, size=int(rng.integers(1, R)), replace=False))
        rest = [w for w in out_w if w not in lw]
        rw = ks[: max(0, nk - trial % 2)] + sorted(set(rest) | set(int(w) for w in rng.choice(out_w, size=int(rng.integers(1, R)), replace=False)))
        hi_r = dt[tr][1] if tl != "i64" else min(dt[tr][1], 1 << 20)      # keep products inside int64
        L = rng.integers(0, dt[tl][1], size=(2,) * len(lw)).astype(dt[tl][0])
        Rr = rng.integers(0, hi_r, size=(2,) * len(rw)).astype(dt[tr][0])
        dev = _native.device()
        Ld, Rd = torch.from_numpy(L.reshape(-1)).to(dev), torch.from_numpy(Rr.reshape(-1)).to(dev)
        Od = torch.empty(1 << R, dtype=torch.int64, device=dev)
        _native.contract2(Ld, [2] * len(lw), lw, Rd, [2] * len(rw), rw, Od, [2] * R, out_w, 1)
        want = O.einsum2(L.astype(np.int64), lw, Rr.astype(np.int64), rw, out_w)
        assert np.array_equal(Od.cpu().numpy().reshape((2,) * R), want)


def _pack_host(t):
    return np.packbits(t.reshape(-1).astype(np.uint8), bitorder="little")


@pytest.mark.parametrize("n", [1, 5, 16, 31, 32, 33, 100, 4096, 70001])
def test_pack_unpack_bits_round_trip(n):
    """Byte <-> bit conversions against numpy.packbits (little-endian bit order inside each 32-bit word)."""
    import torch

    from tensorsat_b200 import _native

    rng = np.random.default_rng(n)
    t = (rng.random(n) < 0.4).astype(np.uint8)
    dev = _native.device()
    src = torch.from_numpy(t).to(dev)
    packed = torch.zeros(_native.packed_nbytes(n), dtype=torch.uint8, device=dev)
    _native.pack_bits(src, packed)
    want = _pack_host(t)
    assert np.array_equal(packed.cpu().numpy()[: len(want)], want)
    back = torch.full((n,), 7, dtype=torch.uint8, device=dev)
    _native.unpack_bits(packed, back)
    assert np.array_equal(back.cpu().numpy(), t)
    other = (rng.random(n) < 0.3).astype(np.uint8)
    acc = torch.from_numpy(other).to(dev)
    _native.unpack_bits(packed, acc, accumulate=True)
    assert np.array_equal(acc.cpu().numpy(), t | other)


@pytest.mark.parametrize("seed", range(8))
def test_contract2_bits_random_lane_masks(seed):
    """k_merge_bits through tsb_contract2_bits: random subsets of the output wires on each operand (every
    in-word / word-level lane pattern class), 0..3 summed wires, against the oracle's boolean einsum."""
    import torch

    from tensorsat_b200 import _native

    rng = np.random.default_rng(100 + seed)
    dev = _native.device()
    for trial in range(6):
        R = int(rng.integers(7, SCALE["big_bits"] - 2))
        nk = int(rng.integers(0, 4))
        ks = list(range(nk))
        out_w = list(range(nk, R + nk))
        # how many of the 7 lowest output wires each operand holds decides the code path: force variety
        low = out_w[-7:]
        pick_l = [w for w in low if rng.random() < (0.15, 0.5, 0.85, 1.0)[trial % 4]]
        pick_r = [w for w in low if rng.random() < (1.0, 0.85, 0.3, 0.6)[trial % 4]]
        hi = out_w[:-7]
        lw = ks + sorted(set(pick_l) | set(int(w) for w in hi if rng.random() < 0.5))
        rw_k = [k for k in ks if rng.random() < 0.7]
        rw = rw_k + sorted((set(out_w) - set(lw)) | set(pick_r) | set(int(w) for w in hi if rng.random() < 0.4))
        L = (rng.random((2,) * len(lw)) < 0.35).astype(np.uint8)
        Rr = (rng.random((2,) * len(rw)) < 0.45).astype(np.uint8)
        bufs = []
        for t in (L, Rr):
            b = torch.zeros(_native.packed_nbytes(t.size), dtype=torch.uint8, device=dev)
            _native.pack_bits(torch.from_numpy(t.reshape(-1)).to(dev), b)
            bufs.append(b)
        ob = torch.zeros(_native.packed_nbytes(1 << R), dtype=torch.uint8, device=dev)
        _native.contract2_bits(bufs[0], lw, bufs[1], rw, ob, out_w, R + nk)
        want = O.contract2(L, lw, Rr, rw, out_w, mode="sat")
        got = np.unpackbits(ob.cpu().numpy(), bitorder="little")[: 1 << R].reshape((2,) * R)
        assert np.array_equal(got, want), (R, nk, lw, rw)


def test_contract2_bits_every_lane_class():
    """All 128 lane masks on one operand (the other holds every wire), on either side."""
    import torch

    from tensorsat_b200 import _native

    rng = np.random.default_rng(77)
    dev = _native.device()
    R = 9
    out_w = list(range(1, R + 1))          # wire 0 is summed
    full = [0] + out_w
    F = (rng.random((2,) * len(full)) < 0.5).astype(np.uint8)
    fb = torch.zeros(_native.packed_nbytes(F.size), dtype=torch.uint8, device=dev)
    _native.pack_bits(torch.from_numpy(F.reshape(-1)).to(dev), fb)
    for mask in range(128):
        low = [out_w[R - 1 - j] for j in range(6, -1, -1) if mask >> j & 1]   # output bit j = wire out_w[R-1-j]
        sw = [0, out_w[0]] + low
        S = (rng.random((2,) * len(sw)) < 0.5).astype(np.uint8)
        sb = torch.zeros(_native.packed_nbytes(S.size), dtype=torch.uint8, device=dev)
        _native.pack_bits(torch.from_numpy(S.reshape(-1)).to(dev), sb)
        for small_first in (True, False):
            ob = torch.zeros(_native.packed_nbytes(1 << R), dtype=torch.uint8, device=dev)
            if small_first:
                _native.contract2_bits(sb, sw, fb, full, ob, out_w, R + 1)
            else:
                _native.contract2_bits(fb, full, sb, sw, ob, out_w, R + 1)
            want = O.contract2(S, sw, F, full, out_w, mode="sat")
            got = np.unpackbits(ob.cpu().numpy(), bitorder="little")[: 1 << R].reshape((2,) * R)
            assert np.array_equal(got, want), (mask, small_first)


@pytest.mark.parametrize("nk", [0, 1])
def test_contract2_bits_one_missing_in_word_lane_takes_the_word_path(nk):
    """An operand that lacks exactly one in-word lane (and is too large for a broadcast twin in a plan) against a
    full-lane one: k_merge_bits_w with an 8-byte fetch and block-wise bit doubling, every missing lane, either side."""
    import torch

    from tensorsat_b200 import _native

    rng = np.random.default_rng(31 + nk)
    dev = _native.device()
    R = 10
    ks = list(range(nk))
    out_w = list(range(nk, R + nk))
    full = ks + out_w
    F = (rng.random((2,) * len(full)) < 0.5).astype(np.uint8)
    fb = torch.zeros(_native.packed_nbytes(F.size), dtype=torch.uint8, device=dev)
    _native.pack_bits(torch.from_numpy(F.reshape(-1)).to(dev), fb)
    for p in range(5):
        sw = ks + [w for j, w in enumerate(out_w) if R - 1 - j != p]   # output bit p = wire out_w[R - 1 - p]
        S = (rng.random((2,) * len(sw)) < 0.5).astype(np.uint8)
        sb = torch.zeros(_native.packed_nbytes(S.size), dtype=torch.uint8, device=dev)
        _native.pack_bits(torch.from_numpy(S.reshape(-1)).to(dev), sb)
        for small_first in (True, False):
            ob = torch.zeros(_native.packed_nbytes(1 << R), dtype=torch.uint8, device=dev)
            _native.trace(True)
            try:
                if small_first:
                    _native.contract2_bits(sb, sw, fb, full, ob, out_w, R + nk)
                else:
                    _native.contract2_bits(fb, full, sb, sw, ob, out_w, R + nk)
                counts = {k: v for k, v in _native.kernel_counts().items() if v}
            finally:
                _native.trace(False)
            assert counts == {"k_merge_bits_w": 1}, counts
            want = O.contract2(S, sw, F, full, out_w, mode="sat")
            got = np.unpackbits(ob.cpu().numpy(), bitorder="little")[: 1 << R].reshape((2,) * R)
            assert np.array_equal(got, want), (p, small_first)


def test_contract2_bits_rejects_what_it_cannot_do():
    import torch

    from tensorsat_b200 import _native

    dev = _native.device()
    b = torch.zeros(64, dtype=torch.uint8, device=dev)
    with pytest.raises(NotImplementedError):   # out_rank < 7
        _native.contract2_bits(b, [0, 1], b, [1, 2], b, [0, 2], 3)
    with pytest.raises(NotImplementedError):   # wire order of the output disagrees with the operands
        _native.contract2_bits(b, list(range(8)), b, list(range(8)), b, list(range(7, -1, -1)), 8)


@pytest.mark.parametrize("k,r", [(9, 0), (13, 0), (14, 2), (12, 5), (17, 1)])
def test_contract2_many_summed_wires(k, r):
    """Dot-product-like steps (many summed wires, few outputs) are split over the summed wires (split-K)."""
    if k + r > SCALE["big_bits"] + 3:
        pytest.skip("too large for the emulated build")
    rng = np.random.default_rng(k * 10 + r)
    kw, ow = list(range(k)), list(range(k, k + r))
    lw = [int(w) for w in rng.permutation(kw + ow[: (r + 1) // 2])]
    rw = [int(w) for w in rng.permutation(kw + ow[(r + 1) // 2:])]
    lhs = _rand01(rng, (2,) * len(lw), 0.02)
    rhs = _rand01(rng, (2,) * len(rw), 0.02)
    got = CudaFinRel.contract2(CudaFinRel(lhs), lw, CudaFinRel(rhs), rw, ow)
    want = O.contract2(lhs, lw, rhs, rw, ow)
    assert np.array_equal(got.tensor, want)
    dense_l, dense_r = _rand01(rng, (2,) * len(lw), 0.5), _rand01(rng, (2,) * len(rw), 0.5)
    gotc = CudaCountRel.contract2(CudaCountRel(dense_l), lw, CudaCountRel(dense_r), rw, ow)
    assert np.array_equal(gotc.tensor, O.contract2(dense_l, lw, dense_r, rw, ow, mode="count"))


def test_contract2_validation_errors():
    a, b = CudaFinRel(np.ones((2, 2), np.uint8)), CudaFinRel(np.ones((2,), np.uint8))
    with pytest.raises(ValueError):
        CudaFinRel.contract2(a, [0], b, [1])  # wrong number of lhs wires
    with pytest.raises(ValueError):
        CudaFinRel.contract2(a, [0, 1], b, [1], [2])  # output wire from nowhere
    with pytest.raises(TypeError):
        CudaFinRel.contract2(a, [0, 1], CudaCountRel(np.ones((2,), np.int64)), [1])
    with pytest.raises(ValueError):
        CudaFinRel(np.full((2,), 2, np.uint8))
    with pytest.raises(ValueError):
        bool(a)
    with pytest.raises(NotImplementedError):
        a[(slice(0, 1), 0)]
    with pytest.raises(ValueError):
        a.rewire([2])


# ------------------------------------------------------------------------------------ permutes
def _resident(cls, arr):
    """A box whose data lives ONLY on the device: a lazy ``rewire`` of it is materialised by the permute kernels
    (FinRel._rewire's transpose view, tensorsat/lang/fin_rel.py:379-382, walked by the next consumer)."""
    box = cls(arr)
    box._physical()
    box._host = None
    return box


def _permute_on_device(box, perm):
    """``box.rewire(perm).device_tensor`` with proof that a permute kernel produced it."""
    from tensorsat_b200 import _native

    _native.trace(True)
    before = _native.launch_count()
    out = box.rewire(perm).device_tensor.cpu().numpy()
    counts = _native.kernel_counts()
    _native.trace(False)
    assert _native.launch_count() > before
    return out, counts


@pytest.mark.parametrize("seed", range(8))
def test_permute_random(seed):
    rng = np.random.default_rng(3000 + seed)
    r = int(rng.integers(9, SCALE["big_bits"] + 1))
    t = _rand01(rng, (2,) * r)
    perm = [int(p) for p in rng.permutation(r)]
    got, counts = _permute_on_device(_resident(CudaFinRel, t), perm)
    assert np.array_equal(got, np.transpose(t, perm))
    assert perm == sorted(perm) or any(k.startswith("k_permute") for k in counts), counts


def test_permute_byte_transposes_in_registers():
    """Permutations that move the lowest axes away (4x4 byte-transpose kernel) incl. full reversal."""
    rng = np.random.default_rng(6)
    seen = {}
    for r in (9, 12, 14, min(17, SCALE["big_bits"])):
        t = rng.integers(0, 2, size=(2,) * r, dtype=np.uint8)
        box = _resident(CudaFinRel, t)
        perms = [list(range(r))[::-1], list(range(2, r)) + [0, 1], [r - 1, r - 2] + list(range(r - 2))]
        perms.append([int(p) for p in rng.permutation(r - 6)] + [r - 1, r - 2, r - 6, r - 5, r - 4, r - 3])
        for perm in perms:
            got, counts = _permute_on_device(box, perm)
            assert np.array_equal(got, np.transpose(t, perm)), (r, perm)
            for k, v in counts.items():
                seen[k] = seen.get(k, 0) + v
    assert seen.get("k_permute_t4", 0) >= 4, seen


def test_permute_keeps_low_axes_and_wide_elements():
    rng = np.random.default_rng(5)
    r = 13
    t = rng.integers(0, 1 << 40, size=(2,) * r, dtype=np.int64)
    box = _resident(CudaCountRel, t)
    for perm in ([1, 0] + list(range(2, r)), list(range(r - 3, -1, -1)) + [r - 2, r - 1], list(range(r))[::-1]):
        got, counts = _permute_on_device(box, perm)
        assert np.array_equal(got, np.transpose(t, perm))
        assert any(k in counts for k in ("k_permute<8>", "k_permute<16>")), counts
    u = _rand01(rng, (2,) * 14)
    ubox = _resident(CudaFinRel, u)
    for keep in (1, 2, 3, 4, 5):  # low axes left in place: gather groups of 2..16 bytes
        perm = [int(p) for p in rng.permutation(14 - keep)] + list(range(14 - keep, 14))
        got, counts = _permute_on_device(ubox, perm)
        assert np.array_equal(got, np.transpose(u, perm))
        # (the random head may leave further low axes in place: the gather group is then wider)
        assert any(f"k_permute<{g}>" in counts for g in (2, 4, 8, 16) if g >= min(16, 1 << keep)), (keep, counts)


@pytest.mark.parametrize("elem_size,keep,kernel", [
    (1, 0, "k_permute<1>"), (1, 1, "k_permute<2>"), (1, 2, "k_permute<4>"), (1, 3, "k_permute<8>"), (1, 4, "k_permute<16>"),
    (4, 0, "k_permute<4>"), (4, 1, "k_permute<8>"), (4, 2, "k_permute<16>"),
    (8, 0, "k_permute<8>"), (8, 1, "k_permute<16>"), (1, -1, "k_permute_t4"),
])
@pytest.mark.parametrize("rank", [14, 16])
def test_permute_every_kernel_variant(elem_size, keep, kernel, rank):
    """tsb_permute through the C-ABI on device buffers: every gather width of k_permute<G> (bytes the
    permutation leaves contiguous, element size included), the 4x4 register-transpose kernel, element sizes
    1 / 4 / 8, ranks >= 14 — each checked against np.transpose and against the kernel the library reports."""
    import torch

    from tensorsat_b200 import _native

    if rank > SCALE["big_bits"] + 1:
        pytest.skip("too large for the emulated build")
    rng = np.random.default_rng(rank * 100 + elem_size * 10 + keep + 1)
    dt = {1: np.uint8, 4: np.int32, 8: np.int64}[elem_size]
    t = rng.integers(0, 200, size=(2,) * rank).astype(dt)
    if keep < 0:      # no low axis stays, and the two lowest source axes leave the four lowest destination axes
        head = [int(p) for p in rng.permutation(rank - 2)]
        perm = head[:3] + [rank - 1, rank - 2] + head[3:]
    else:
        moved = rank - keep
        while True:   # the lowest moved axis must really move (else one more byte group stays contiguous) ...
            head = [int(p) for p in rng.permutation(moved)]
            if head[-1] != moved - 1 and (elem_size > 1 or keep > 0 or moved - 1 in head[-4:] or moved - 2 in head[-4:]):
                break  # ... and with nothing kept, byte tensors stay off the register-transpose variant
        perm = head + list(range(moved, rank))
    src = torch.from_numpy(t.reshape(-1).copy()).to(_native.device())
    dst = torch.empty_like(src)
    _native.trace(True)
    _native.permute(src, dst, (2,) * rank, perm, elem_size)
    counts = _native.kernel_counts()
    _native.trace(False)
    assert np.array_equal(dst.cpu().numpy().reshape((2,) * rank), np.transpose(t, perm)), perm
    assert counts == {kernel: 1}, (perm, counts)


# ------------------------------------------------------------------------------------ whole networks
def _network_and_leaves(diagram):
    flat = diagram.flatten()
    return flat, Network.from_diagram(flat)


def _oracle_tree(flat, tree, mode="sat"):
    leaves = [b.tensor for b in flat.boxes]
    return O.execute_tree(tree.network.inputs, tree.network.output, leaves, tree.ssa_path,
                          tree.sliced_wires, mode=mode, size_dict=tree.network.size_dict)


@pytest.mark.parametrize("k,n,m,seed,target,batch", [(3, 14, 50, 0, 10, 6), (3, 15, 55, 1, 11, 9), (4, 14, 110, 2, 30, 24),
                                                    (3, 16, 62, 3, 12, 24), (3, 15, 58, 4, 9, 3)])
def test_bit_packed_plans_give_the_same_solution_tensors(monkeypatch, k, n, m, seed, target, batch):
    """All-solutions tensors (2^n entries, zeros and ones) through plans whose large steps run bit-packed
    (threshold lowered by the test knob so that these sizes qualify): pre / per-slice / gathered / post paths,
    conversions on both sides, against brute force and against the same plan without packing."""
    monkeypatch.setenv("TSB_PACK_MIN_LOG2", "0")
    cnf = CNFInstance.random(k, n, m, rng=seed)
    flat, net = _network_and_leaves(cnf.diagram())
    tree = find_tree_elimination(net, trials=2, seed=seed)
    if target < 30:   # output wires cannot be sliced: slice a few internal wires instead
        internal = [w for w in range(n, net.num_wires)]
        rs = np.random.default_rng(seed)
        tree = tree.with_sliced(sorted(int(w) for w in rs.choice(internal, size=2 + seed % 2, replace=False)))
    want = O.cnf_brute_force(n, cnf.clauses).reshape((2,) * n)
    assert 0 < np.count_nonzero(want) < want.size
    packed = TreeContraction(CudaFinRel, flat.wiring, tree=tree, batch_log2=batch)
    plain = TreeContraction(CudaFinRel, flat.wiring, tree=tree, batch_log2=batch, bit_packing=False)
    assert packed.plan().info["n_packed_steps"] > 0 and plain.plan().info["n_packed_steps"] == 0
    assert np.array_equal(packed.contract(flat).tensor, want)
    assert np.array_equal(plain.contract(flat).tensor, want)
    # SAT decision with all inputs unknown on the same instance (scalar root, accumulated over slices)
    dd, dnet = _network_and_leaves(cnf.inputs(None) >> cnf.diagram())
    dtree = find_tree_elimination(dnet, trials=2, seed=seed)
    if dtree.width() > target:
        dtree = dtree.slice_to(target)
    dec = TreeContraction(CudaFinRel, dd.wiring, tree=dtree, batch_log2=batch)
    assert bool(dec.contract(dd)) == bool(want.any())
    if dec.num_slices >= 2:   # a partial slice range and its complement OR to the full answer
        h = dec.num_slices // 2
        a = dec.contract(dd, slice_range=(0, h)).tensor | dec.contract(dd, slice_range=(h, dec.num_slices)).tensor
        assert bool(a) == bool(want.any())


@pytest.mark.parametrize("n,m,seed,batch,n_out,fuse_k,pair", [
    (22, 90, 6, 13, 6, 2, 2), (22, 92, 2, 11, 9, 2, 2), (24, 100, 6, 15, 5, 2, 2), (24, 96, 4, 12, 7, 2, 2),
    (22, 90, 6, 13, 6, 3, 2), (22, 92, 2, 11, 9, 3, 2), (22, 92, 2, 11, 9, -1, 2),
    (22, 90, 6, 13, 6, 2, 1), (22, 92, 2, 11, 9, 2, 0), (24, 100, 6, 15, 5, 1, 1), (24, 96, 4, 12, 7, 2, -1),
    (24, 96, 4, 12, 7, 2, 3), (26, 106, 5, 14, 5, 2, 3), (22, 90, 3, 14, 5, 2, 2)])
def test_bit_packed_sliced_plans_partial_outputs(monkeypatch, n, m, seed, batch, n_out, fuse_k, pair):
    """Marginals over a few variables (2^n_out entries, zeros and ones) through sliced plans whose pre, per-slice,
    gathered and post steps all run bit-packed, with broadcast twins: equal to brute force, to the byte plan, and a
    partial slice range plus its complement OR to the full answer."""
    import torch

    from tensorsat_b200 import _native

    monkeypatch.setenv("TSB_PACK_MIN_LOG2", "0")
    monkeypatch.setenv("TSB_FUSE_MIN_LOG2", "0")       # fuse consecutive packed steps whatever their size ...
    monkeypatch.setenv("TSB_FUSE_MAX_K", str(fuse_k))  # ... up to this many summed wires per launch (-1: never)
    monkeypatch.setenv("TSB_FUSE_PAIR", str(pair))     # pair wires of k_merge_bits_f (-1: k_merge_bits_n only; 2, 3: private streaming)
    monkeypatch.setenv("TSB_BITSF_MIN_ITEMS_LOG2", "1")  # ... also on steps this small
    tile = seed in (2, 4) and pair < 3                 # tile wires chosen by operand footprint (TSB_FUSE_TILE) ...
    monkeypatch.setenv("TSB_FUSE_TILE", "1" if tile else "0")
    monkeypatch.setenv("TSB_FUSE_LANES", "0" if (seed, pair) == (6, 1) or tile else "1")   # ... or lane wires by wavefronts
    cnf = CNFInstance.random(3, n, m, rng=seed)
    flat, net0 = _network_and_leaves(cnf.inputs(None) >> cnf.diagram())
    outs = tuple(int(v) for v in np.linspace(0, n - 1, n_out).astype(int))
    net = Network(net0.inputs, outs, net0.size_dict)
    tree = find_tree_elimination(net, trials=2, seed=seed)
    tree = tree.slice_to(max(tree.width() - 3, n_out + 2))
    assert tree.num_slices >= 2
    brute = O.cnf_brute_force(n, cnf.clauses).reshape((2,) * n)
    want = brute.any(axis=tuple(v for v in range(n) if v not in outs)).astype(np.uint8)
    assert 0 < np.count_nonzero(want) < want.size
    leaves = [_native.upload(b.tensor.astype(np.uint8)) for b in flat.boxes]
    got = {}
    for flags in (0, _native.PLAN_NO_PACK):
        plan = _native.Plan(net.inputs, outs, net.size_dict, tree.ssa_path, tree.sliced_wires, mode=_native.MODE_SAT,
                            batch_log2=batch, flags=flags)
        assert (plan.info["n_packed_steps"] > 0) == (flags == 0)
        assert (plan.info["n_fused_steps"] > 0) == (flags == 0 and fuse_k >= 0)
        if flags == 0 and 0 <= fuse_k <= 2:   # the factored kernel takes every fused step with <= 2 summed wires
            fused = [o for o in plan.describe()["ops"] if o["fused"] > 0]
            assert fused and all(o["factored"] == (pair >= 0) for o in fused)
        if flags == 0 and (pair == 3 or (n, seed) == (22, 3)):   # steps with two / three output wires private to one operand
            assert max(o["n_pair"] for o in plan.describe()["ops"]) >= (3 if n == 24 else 2)
        if flags == 0 and (n, m, seed) == (24, 100, 6):   # this plan gathers a per-slice result that stays bit-packed
            assert any(o["per_slice"] and o["out_gathered"] and o["out_packed"] for o in plan.describe()["ops"])
        res = _native.empty(plan.info["out_numel"], _native.MODE_SAT)
        plan.execute(leaves, res)
        got[flags] = res.cpu().numpy().reshape(want.shape).copy()
        h = plan.info["n_slices"] // 2
        a, b = torch.zeros_like(res), torch.zeros_like(res)
        plan.execute(leaves, a, 0, h)
        plan.execute(leaves, b, h, plan.info["n_slices"])
        assert np.array_equal((a | b).cpu().numpy().reshape(want.shape), want)
    assert np.array_equal(got[0], want) and np.array_equal(got[_native.PLAN_NO_PACK], want)


@pytest.mark.parametrize("na,nb,nc,n_c,pair,want_pairs", [
    (3, 2, 9, 9, 3, (3, 2)), (3, 2, 10, 10, 3, (3, 2)), (3, 1, 9, 7, 3, (3, 1)), (3, 2, 9, 9, 3, (3, 0)), (2, 2, 9, 8, 2, (2, 2)),
    (3, 2, 9, 9, 1, (1, 0))])
def test_private_streaming_fused_kernel(monkeypatch, na, nb, nc, n_c, pair, want_pairs):
    """k_merge_bits_f with pair wires private to one operand (slot 0 streamed over 4 / 8 combinations) and the
    two-sided form (slot 1 holds private wires too): out[a.., b.., c..] = OR_{k1,k2} A[k1,k2,a..,c..] B[k2,k1,b..,c..]
    C[k1,c..] as one fused launch, against the integer einsum + sign of FinRel._contract2 (fin_rel.py:276-279)."""
    from tensorsat_b200 import _native

    monkeypatch.setenv("TSB_PACK_MIN_LOG2", "0")
    monkeypatch.setenv("TSB_FUSE_MIN_LOG2", "0")
    monkeypatch.setenv("TSB_BITSF_MIN_ITEMS_LOG2", "1")
    monkeypatch.setenv("TSB_FUSE_PAIR", str(pair))
    monkeypatch.setenv("TSB_BITSF_TWO_SIDED", "0" if want_pairs == (3, 0) else "2")
    rng = np.random.default_rng(na * 100 + nb * 10 + nc)
    aw = list(range(2, 2 + na))
    bw = list(range(2 + na, 2 + na + nb))
    cw = list(range(2 + na + nb, 2 + na + nb + nc))
    inputs = [[0, 1] + aw + cw, [1, 0] + bw + cw, [0] + cw[:n_c]]
    hi = aw + bw
    rng.shuffle(hi)
    out = [int(w) for w in hi] + cw
    leaves = [(rng.random((2,) * len(ws)) < 0.62).astype(np.uint8) for ws in inputs]
    plan = _native.Plan(inputs, out, [2] * (2 + na + nb + nc), [(0, 1), (3, 2)], [], mode=_native.MODE_SAT)
    fused = [o for o in plan.describe()["ops"] if o["kind"] == "bits" and o["fused"] > 0]
    assert len(fused) == 1 and (fused[0]["n_pair"], fused[0]["n_pair2"]) == want_pairs, [(o["n_pair"], o["n_pair2"]) for o in fused]
    res = _native.empty(plan.info["out_numel"], _native.MODE_SAT)
    letters = "abcdefghijklmnopqrstuvwxyz"
    eq = ",".join("".join(letters[w] for w in ws) for ws in inputs) + "->" + "".join(letters[w] for w in out)
    want = (np.einsum(eq, *[x.astype(np.int64) for x in leaves]) > 0).astype(np.uint8)
    assert 0.05 < want.mean() < 0.95
    dev = [_native.upload(x) for x in leaves]
    for _ in range(2):   # direct launches, then the captured graph
        res.zero_()
        plan.execute(dev, res)
        assert np.array_equal(res.cpu().numpy().reshape(want.shape), want)
    plan.close()


def test_box_from_wiring():
    """Box.from_wiring (tensorsat/diagrams.py:750-775): spiders per wire, their product, ports in slot order then out."""
    from tensorsat_b200.diagrams import Wiring

    bit = FinSet(2)
    wiring = Wiring(wire_types=[bit, bit, FinSet(3)], slot_wires_list=[(0, 1), (1, 2)], out_wires=(0, 2, 2))
    box = CudaFinRel.from_wiring(wiring)
    assert box.shape == (bit, bit, bit, FinSet(3), bit, FinSet(3), FinSet(3))
    want = np.zeros((2, 2, 2, 3, 2, 3, 3), dtype=np.uint8)
    for a in range(2):
        for b in range(2):
            for c in range(3):
                want[a, b, b, c, a, c, c] = 1
    assert np.array_equal(box.tensor, want)


def test_assignment_search_on_bit_packed_sliced_plan(monkeypatch):
    """The reference's assignment search (notebooks/02-random-ksat.ipynb:707-722) on one sliced, bit-packed
    plan whose leaves change between executions: every decision must agree with brute force, and the model
    found must satisfy the formula."""
    monkeypatch.setenv("TSB_PACK_MIN_LOG2", "0")
    n, m, seed = 24, 100, 6
    cnf = CNFInstance.random(3, n, m, rng=seed)
    wiring = cnf.contraction_wiring()
    tree = find_tree_elimination(Network.from_wiring(wiring), trials=2, seed=0)
    tree = tree.slice_to(tree.width() - 2)
    c = TreeContraction(CudaFinRel, wiring, tree=tree, batch_log2=13)
    kinds = {(o["kind"], o["per_slice"], o["post"]) for o in c.plan().describe()["ops"]}
    assert {("bits", 1, 0), ("bits", 0, 1), ("unpack", 1, 0)} <= kinds and c.num_slices == 4
    brute = O.cnf_brute_force(n, cnf.clauses).reshape((2,) * n)
    assert brute.any()
    prefix = ""
    for _ in range(n):
        for bit in "01":
            spec = prefix + bit + "?" * (n - len(prefix) - 1)
            got = bool(c.contract((cnf.inputs(spec) >> cnf.diagram()).flatten()))
            want = bool(brute[tuple(int(ch) for ch in prefix + bit)].any())
            assert got == want, spec
            if got:
                prefix += bit
                break
    assert brute[tuple(int(ch) for ch in prefix)] == 1


def test_kissat_verdicts():
    """examples/test_sat_solver.py:22-190 through the DIMACS parser and the plan executor."""
    import os

    from tests.conftest import GOLDEN

    g = load_golden("networks.json")["kissat"]
    checked = 0
    for rec in g:
        if rec["log2_max_size"] > SCALE["big_bits"] + 3:
            continue
        with open(os.path.join(GOLDEN, "kissat", rec["name"] + ".cnf")) as f:
            cnf = CNFInstance.from_dimacs(f.read())
        diagram = (cnf.inputs(None) >> cnf.diagram()).flatten()
        contraction = TreeContraction(CudaFinRel, diagram.wiring, ssa_path=rec.get("ssa_path"), trials=2)
        assert bool(contraction.contract(diagram)) == rec["expected"], rec["name"]
        checked += 1
    assert checked >= 39


def test_random_ksat_solution_tensors_and_counts():
    for rec in load_golden("networks.json")["random"]:
        cnf = CNFInstance.random(rec["k"], rec["n"], rec["m"], rng=rec["seed"])
        diagram = cnf.diagram().flatten()
        tc = TreeContraction(CudaFinRel, diagram.wiring, ssa_path=rec["ssa_path"])
        sol = tc.contract(diagram)
        got = sorted("".join(map(str, p)) for p in sol.to_set())
        assert got == rec["solutions"]
        brute = O.cnf_brute_force(cnf.num_vars, cnf.clauses).reshape((2,) * cnf.num_vars)
        assert np.array_equal(sol.tensor, brute)
        # decisions with partial inputs (the reference's assignment search, 02-random-ksat.ipynb:707-722)
        for dec in rec["decisions"]:
            dd = (cnf.inputs(dec["inputs"]) >> cnf.diagram()).flatten()
            c2 = TreeContraction(CudaFinRel, dd.wiring, ssa_path=dec["ssa_path"])
            assert bool(c2.contract(dd)) == dec["sat"]
        # #SAT: same network, counting semiring, all inputs unknown
        dd = (cnf.inputs(None) >> cnf.diagram()).flatten()
        cc = TreeContraction(CudaCountRel, dd.wiring, trials=2, wire_is_free=[w < cnf.num_vars for w in dd.wiring.wires])
        assert int(cc.contract(dd)) == len(rec["solutions"]) == int(brute.sum())


def test_circuits_truth_tables():
    diagrams = {"half_adder": bc.half_adder, "full_adder": bc.full_adder, "rc_adder(2)": bc.rc_adder(2),
                "wallace_multiplier(2)": bc.wallace_multiplier(2), "wallace_multiplier(3)": bc.wallace_multiplier(3)}
    for rec in load_golden("networks.json")["circuits"]:
        d = diagrams[rec["name"]].flatten()
        res = TreeContraction(CudaFinRel, d.wiring, ssa_path=rec["ssa_path"]).contract(d)
        assert sorted("".join(map(str, p)) for p in res.to_set()) == rec["points"]
    # multiplier semantics: a * b == p on every row
    d = bc.wallace_multiplier(3).flatten()
    res = TreeContraction(CudaFinRel, d.wiring, trials=2).contract(d)
    for pt in res.to_set():
        a = sum(b << i for i, b in enumerate(pt[0:3]))
        b_ = sum(b << i for i, b in enumerate(pt[3:6]))
        p = sum(b << i for i, b in enumerate(pt[6:]))
        assert a * b_ == p


def test_simple_contraction_matches_reference():
    for rec in load_golden("networks.json")["simple"]:
        if "name" in rec:
            d = bc.full_adder
            res = SimpleContraction(CudaFinRel, d.wiring, rec["linear_path"]).contract(d)
            assert sorted("".join(map(str, p)) for p in res.to_set()) == rec["points"]
        else:
            cnf = CNFInstance.random(rec["k"], rec["n"], rec["m"], rng=rec["seed"])
            d = cnf.diagram().flatten()
            res = SimpleContraction(CudaFinRel, d.wiring, rec["linear_path"]).contract(d)
            assert sorted("".join(map(str, p)) for p in res.to_set()) == rec["solutions"]


@pytest.mark.parametrize("n,m,seed,target", [(14, 58, 3, 7), (16, 66, 1, 8), (12, 50, 4, 5)])
def test_sliced_matches_unsliced_and_oracle(n, m, seed, target):
    cnf = CNFInstance.random(3, n, m, rng=seed)
    diagram = (cnf.inputs(None) >> cnf.diagram()).flatten()
    flat, net = _network_and_leaves(diagram)
    tree = find_tree_elimination(net, trials=4, seed=seed)
    sliced = tree.slice_to(target)
    assert sliced.num_slices > 1
    brute = O.cnf_brute_force(n, cnf.clauses)
    for box_class, mode in ((CudaFinRel, "sat"), (CudaCountRel, "count")):
        free = [w < n for w in flat.wiring.wires]
        full = TreeContraction(box_class, flat.wiring, tree=tree, wire_is_free=free).contract(flat)
        part = TreeContraction(box_class, flat.wiring, tree=sliced, wire_is_free=free, batch_log2=4).contract(flat)
        lazy = TreeContraction(box_class, flat.wiring, tree=sliced, wire_is_free=free).contract(flat)
        want = _oracle_tree(flat, sliced, mode)
        assert np.array_equal(full.tensor, want) and np.array_equal(part.tensor, want) and np.array_equal(lazy.tensor, want)
        if mode == "count":
            assert int(full) == int(brute.sum())
        else:
            assert bool(full) == bool(brute.any())
        # slice ranges (what the ranks of a multi-GPU job execute) combine to the whole
        for bl in (-1, 4):
            tc = TreeContraction(box_class, flat.wiring, tree=sliced, wire_is_free=free, batch_log2=bl)
            k = sliced.num_slices // 3
            parts = [tc.contract(flat, slice_range=r).tensor for r in ((0, k), (k, k), (k, sliced.num_slices))]
            comb = parts[0] + parts[1] + parts[2] if mode == "count" else parts[0] | parts[1] | parts[2]
            assert np.array_equal(comb, want)


def test_sliced_output_wires_are_stacked():
    """Sliced wires that are output wires: per-slice results are stacked (Box.stack, diagrams.py:883-896)."""
    cnf = CNFInstance.random(3, 8, 20, rng=2)
    diagram = cnf.diagram().flatten()
    flat, net = _network_and_leaves(diagram)
    tree = find_tree_elimination(net, trials=2, seed=0)
    want = O.cnf_brute_force(8, cnf.clauses).reshape((2,) * 8)
    for sl in ((2,), (5, 0), (1, 7, 3)):
        t = tree.with_sliced(sl)
        got = TreeContraction(CudaFinRel, flat.wiring, tree=t, batch_log2=3).contract(flat)
        assert np.array_equal(got.tensor, want)
        gotc = TreeContraction(CudaCountRel, flat.wiring, tree=t, batch_log2=3,
                               wire_is_free=[w < 8 for w in flat.wiring.wires]).contract(flat)
        assert np.array_equal(gotc.tensor, want.astype(np.int64))
        with pytest.raises(OverflowError):  # 74 summed wires, none declared as determined
            TreeContraction(CudaCountRel, flat.wiring, tree=t).contract(flat)


def test_generic_and_fast_plans_agree():
    cnf = CNFInstance.random(3, 18, 60, rng=9)
    diagram = (cnf.inputs("0?1" + "?" * 15) >> cnf.diagram()).flatten()
    flat, net = _network_and_leaves(diagram)
    tree = find_tree_elimination(net, trials=2, seed=1)
    free = [w < 18 for w in flat.wiring.wires]
    a = TreeContraction(CudaCountRel, flat.wiring, tree=tree, wire_is_free=free).contract(flat)
    b = TreeContraction(CudaCountRel, flat.wiring, tree=tree, wire_is_free=free, force_generic=True).contract(flat)
    assert int(a) == int(b) == int(_oracle_tree(flat, tree, "count"))


def test_plan_reuse_with_changed_leaves():
    """One contraction object, many input assignments (notebooks/01-introduction.ipynb:1824-1849)."""
    cnf = CNFInstance.random(3, 3, 6, rng=3)  # golden G5
    wiring = cnf.contraction_wiring()
    contraction = TreeContraction(CudaFinRel, wiring, trials=2)
    trace = {}
    for spec in ("0??", "00?", "01?", "010", "011", "1??", "111"):
        d = (cnf.inputs(spec) >> cnf.diagram()).flatten()
        trace[spec] = bool(contraction.contract(d))
    assert trace == {"0??": True, "00?": False, "01?": True, "010": False, "011": True, "1??": True, "111": True}


def test_arena_is_reused_across_executions():
    from tensorsat_b200 import _native

    cnf = CNFInstance.random(3, 12, 50, rng=4)
    d = (cnf.inputs(None) >> cnf.diagram()).flatten()
    base = _native.arena_stats()
    c = TreeContraction(CudaFinRel, d.wiring, trials=2)
    c.contract(d)
    c.contract(d)   # (on the GPU the second execution is captured into a CUDA graph that owns its arena block)
    before = _native.arena_stats()
    for _ in range(3):
        c.contract(d)
    after = _native.arena_stats()
    assert after["device_allocs"] == before["device_allocs"]
    replays = c.plan().describe()["graph_replays"]
    assert replays >= 3 or after["cache_hits"] >= before["cache_hits"] + 3
    assert after["in_use_bytes"] == before["in_use_bytes"]
    c.plan().close()   # a plan returns what its graphs held
    assert _native.arena_stats()["in_use_bytes"] == base["in_use_bytes"]


# ------------------------------------------------------------------------------------ autoray backend
def test_arena_orders_reuse_across_streams():
    """A block released on one stream and handed to another is ordered by an event (stream-aware arena); the
    stream-less forms keep working; the bookkeeping survives threads."""
    import ctypes as C
    import threading

    from tensorsat_b200 import _native

    lib, ctx = _native.lib(), _native.ctx()
    if _native._device == "cuda":
        import torch

        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        a, b = C.c_void_p(s1.cuda_stream), C.c_void_p(s2.cuda_stream)
    else:
        a, b = C.c_void_p(16), C.c_void_p(32)
    n = 3 << 20
    p, q = C.c_void_p(), C.c_void_p()
    _native._check(lib.tsb_arena_alloc_async(ctx, C.c_size_t(n), a, C.byref(p)))
    w0 = _native.arena_stats()["cross_stream_waits"]
    _native._check(lib.tsb_arena_release_async(ctx, p, a))
    _native._check(lib.tsb_arena_alloc_async(ctx, C.c_size_t(n), a, C.byref(q)))     # same stream: no wait needed
    assert q.value == p.value and _native.arena_stats()["cross_stream_waits"] == w0
    _native._check(lib.tsb_arena_release_async(ctx, q, a))
    _native._check(lib.tsb_arena_alloc_async(ctx, C.c_size_t(n), b, C.byref(q)))     # other stream: ordered by the event
    assert q.value == p.value and _native.arena_stats()["cross_stream_waits"] == w0 + 1
    _native._check(lib.tsb_arena_release_async(ctx, q, b))
    _native._check(lib.tsb_arena_alloc(ctx, C.c_size_t(n), C.byref(q)))              # stream-less: host-side wait
    assert q.value == p.value and _native.arena_stats()["cross_stream_waits"] == w0 + 2
    _native._check(lib.tsb_arena_release(ctx, q))
    with pytest.raises(ValueError):
        _native._check(lib.tsb_arena_release(ctx, q))    # double release is refused, not corrupted
    errors = []

    def worker():
        try:
            for _ in range(200):
                r = C.c_void_p()
                _native._check(lib.tsb_arena_alloc_async(ctx, C.c_size_t(1 << 12), a, C.byref(r)))
                _native._check(lib.tsb_arena_release_async(ctx, r, a))
        except Exception as exc:  # pragma: no cover
            errors.append(exc)

    threads = [threading.Thread(target=worker) for _ in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    st = _native.arena_stats()
    assert not errors and st["in_use_bytes"] <= st["reserved_bytes"]


def test_autoray_backend_functions():
    """einsum / tensordot / transpose / reshape / stack / shape with the reference's signatures (_autoray.py:31-80)."""
    from tensorsat_b200 import autoray_backend as ab

    rng = np.random.default_rng(3)
    a, b = _rand01(rng, (2, 2, 2, 2)), _rand01(rng, (2, 2, 2))
    A, B = CudaFinRel(a), CudaFinRel(b)
    assert ab.shape(A) == (2, 2, 2, 2)
    assert np.array_equal(ab.einsum("abcd,cde->abe", A, B).tensor, np.sign(np.einsum("abcd,cde->abe", a, b)).astype(np.uint8))
    assert np.array_equal(ab.einsum("abcd->ca", A).tensor, np.sign(np.einsum("abcd->ca", a.astype(np.int64))).astype(np.uint8))
    got = ab.tensordot(A, B, axes=([2, 3], [0, 1]))
    assert np.array_equal(got.tensor, np.sign(np.tensordot(a.astype(np.int64), b, axes=([2, 3], [0, 1]))).astype(np.uint8))
    got = ab.tensordot(A, B, axes=1)
    assert np.array_equal(got.tensor, np.sign(np.tensordot(a.astype(np.int64), b, axes=1)).astype(np.uint8))
    assert np.array_equal(ab.transpose(A, (3, 1, 0, 2)).tensor, np.transpose(a, (3, 1, 0, 2)))
    assert np.array_equal(ab.transpose(B).tensor, np.transpose(b))
    assert np.array_equal(ab.reshape(ab.transpose(A, (1, 0, 2, 3)), (4, 4)).tensor, np.transpose(a, (1, 0, 2, 3)).reshape(4, 4))
    assert np.array_equal(ab.reshape(A, (-1, 2)).tensor, a.reshape(-1, 2))
    assert np.array_equal(ab.stack([B, B, B], 1).tensor, np.stack([b, b, b], 1))
    C = CudaCountRel(a)
    assert np.array_equal(ab.tensordot(C, C, axes=2).tensor, np.tensordot(a.astype(np.int64), a.astype(np.int64), axes=2))
    # device-resident operands: transpose is lazy (like the reference's view) and is materialised on the device,
    # by the permute kernel when read out, or folded into the operand's wire order when contracted
    from tensorsat_b200 import _native

    big = _rand01(rng, (2,) * 12)
    perm = [int(p) for p in rng.permutation(12)]
    G = _resident(CudaFinRel, big)
    _native.trace(True)
    T = ab.transpose(G, perm)
    assert _native.kernel_counts() == {}
    assert np.array_equal(T.device_tensor.cpu().numpy(), np.transpose(big, perm))
    assert any(k.startswith("k_permute") for k in _native.kernel_counts())
    T2 = ab.transpose(_resident(CudaFinRel, big), perm)
    eq = "abcdefghijkl,lkjihg->abcdef"
    want = np.sign(np.einsum(eq, np.transpose(big, perm).astype(np.int64), b.reshape(2, 2, 2, 1, 1, 1) * np.ones((2,) * 6, np.int64)))
    got = ab.einsum(eq, T2, CudaFinRel((b.reshape(2, 2, 2, 1, 1, 1) * np.ones((2,) * 6, np.uint8)).astype(np.uint8)))
    _native.trace(False)
    assert np.array_equal(got.tensor, want.astype(np.uint8))
    with pytest.raises(TypeError):
        ab.shape(3)
    with pytest.raises(ValueError):
        ab.reshape(A, (3, 5))
    with pytest.raises(TypeError):
        ab.einsum("ab,bc->ac", CudaFinRel(a[0, 0]), CudaCountRel(a[0, 0]))


def test_box_protocol_misc():
    """prod / union / spider validation / equality / hash / inspectors (diagrams.py:719-996, fin_rel.py:395-493)."""
    xor_, and_, bit = bc.xor_, bc.and_, bc.bit
    p = CudaFinRel.prod([bc.bit_0, bc.bit_1, bc.bit_unk])           # G10
    assert p.to_set() == {(0, 1, 0), (0, 1, 1)}
    assert CudaFinRel.union([and_, xor_]).to_set() == and_.to_set() | xor_.to_set()
    assert xor_.rewire([0, 2, 1]) == xor_ and hash(xor_.rewire([0, 2, 1])) == hash(xor_)
    assert xor_.is_function_graph([0, 1]) and xor_.to_mapping([0, 1])[(1, 0)] == (1,)
    assert not bc.bit_unk.is_function_graph([])
    with pytest.raises(ValueError):
        CudaFinRel.spider(bit, 0)
    with pytest.raises(ValueError):
        CudaFinRel.union([])
    with pytest.raises(ValueError):
        CudaFinRel.union([and_, bc.not_])
    with pytest.raises(ValueError):
        CudaFinRel.stack([and_], 4)
    with pytest.raises(TypeError):
        CudaFinRel.prod([and_, CudaCountRel.from_rel(and_)])
    trit = FinSet(3)
    sp = CudaFinRel.spider(trit, 2)
    assert sp.tensor_shape == (3, 3) and sp.to_set() == {(0, 0), (1, 1), (2, 2)}
    with pytest.raises(TypeError):
        type("Sub", (CudaFinRel,), {})


def test_factoring_with_the_wallace_multiplier():
    """BASELINE config 5 at a contractible size: fix the product bits of a 3x3 / 4x4 Wallace multiplier and
    read the factor pairs off the relation over the input bits."""
    from tensorsat_b200.diagrams import DiagramBuilder

    for nbits, N in ((3, 35), (4, 143), (4, 165)):
        mult = bc.wallace_multiplier(nbits)
        n_out = mult.num_ports - 2 * nbits
        builder = DiagramBuilder()
        ins = builder.add_inputs(bc.bits(2 * nbits))
        outs = mult @ builder[ins]
        for i, w in enumerate(outs):
            ((bc.bit_1 if (N >> i) & 1 else bc.bit_0) @ builder[w])
        d = builder.diagram().flatten()
        res = TreeContraction(CudaFinRel, d.wiring, trials=2).contract(d)
        if nbits == 3:  # repeated output wires: contracted onto distinct wires, then rewired
            builder.add_outputs(ins[:2])
            d2 = builder.diagram().flatten()
            res2 = TreeContraction(CudaFinRel, d2.wiring, trials=2).contract(d2)
            assert res2.to_set() == {pt + pt[:2] for pt in res.to_set()}
        pairs = set()
        for pt in res.to_set():
            a = sum(b << i for i, b in enumerate(pt[:nbits]))
            b_ = sum(b << i for i, b in enumerate(pt[nbits:]))
            pairs.add((a, b_))
        want = {(a, b_) for a in range(1 << nbits) for b_ in range(1 << nbits) if a * b_ == N}
        assert pairs == want and want, (nbits, N, n_out)
        cnt = TreeContraction(CudaCountRel, d.wiring, trials=2, wire_is_free=[w < 2 * nbits for w in d.wiring.wires]).contract(d)
        assert int(cnt.tensor.sum()) == len(want)


def test_contract_from_host_matches_device_leaves():
    """The end-to-end entry point on host buffers (what bench.py's e2e figure times): uint8 arrays take the
    one-pass packing, wider dtypes and plain lists the checked path; both equal the device-leaf call, and the
    staging buffers are reused across calls with different leaf values."""
    n = 12
    cnf = CNFInstance.random(3, n, 40, rng=5)
    brute = O.cnf_brute_force(n, cnf.clauses).reshape((2,) * n)
    wiring = cnf.contraction_wiring()
    c = TreeContraction(CudaFinRel, wiring, tree=find_tree_elimination(Network.from_wiring(wiring), trials=2, seed=0))
    for spec in ("?" * n, "1" + "?" * (n - 1), "01" + "?" * (n - 2)):
        flat = (cnf.inputs(spec) >> cnf.diagram()).flatten()
        want = bool(brute[tuple(int(ch) for ch in spec if ch != "?")].any())
        leaves = [b.tensor for b in flat.boxes]
        assert bool(c.contract(flat)) == want
        assert bool(c.contract_from_host(leaves)) == want
        assert bool(c.contract_from_host([a.astype(np.int64) for a in leaves])) == want
        assert bool(c.contract_from_host([a.tolist() for a in leaves])) == want
    with pytest.raises(ValueError):
        c.contract_from_host(leaves[:-1])
    # the C-ABI entry point behind it (tsb_plan_execute_host): sliced, counting semiring, and its size checks
    tree = find_tree_elimination(Network.from_wiring(wiring), trials=2, seed=0)
    cc = TreeContraction(CudaCountRel, wiring, tree=tree.slice_to(max(tree.width() - 2, 2)),
                         wire_is_free=[w < n for w in wiring.wires])
    flat = (cnf.inputs("?" * n) >> cnf.diagram()).flatten()
    leaves = [b.tensor for b in flat.boxes]
    assert cc.num_slices >= 2 and int(cc.contract_from_host(leaves)) == int(brute.sum())
    h = cc.num_slices // 2
    assert int(cc.contract_from_host(leaves, (0, h))) + int(cc.contract_from_host(leaves, (h, cc.num_slices))) == int(brute.sum())
    blob = np.concatenate([a.reshape(-1) for a in leaves]).astype(np.uint8)
    out = np.zeros(1, dtype=np.int64)
    plan = cc.plan()
    plan.execute_host(blob.ctypes.data, blob.size, out.ctypes.data, out.nbytes)      # pageable host memory works too
    assert int(out[0]) == int(brute.sum())
    with pytest.raises(ValueError):
        plan.execute_host(blob.ctypes.data, blob.size - 1, out.ctypes.data, out.nbytes)
    with pytest.raises(ValueError):
        plan.execute_host(blob.ctypes.data, blob.size, out.ctypes.data, 4)


def test_contract_from_host_with_repeated_output_wires():
    """Repeated output wires (a diagonal of the result, simple.py:287-289): the host entry point returns the same
    tensor as contract(), which rewires the plan's distinct-output result."""
    from tensorsat_b200.diagrams import Wiring

    rng = np.random.default_rng(11)
    wiring = Wiring(wire_types=[FinSet(2)] * 4, slot_wires_list=[[0, 1], [1, 2, 3], [3, 0]], out_wires=[0, 0, 2])
    leaves = [_rand01(rng, (2, 2), 0.7), _rand01(rng, (2, 2, 2), 0.7), _rand01(rng, (2, 2), 0.7)]
    net = Network(((0, 1), (1, 2, 3), (3, 0)), (0, 2), (2, 2, 2, 2))
    c = TreeContraction(CudaFinRel, wiring, ssa_path=[(0, 1), (3, 2)])
    got = c.contract_from_host(leaves)
    want2 = (np.einsum("ab,bcd,da->ac", *[x.astype(np.int64) for x in leaves]) > 0).astype(np.uint8)
    want = np.zeros((2, 2, 2), np.uint8)
    for a in range(2):
        want[a, a, :] = want2[a, :]
    assert got.shape == (2, 2, 2) and np.array_equal(got, want)
    assert net.num_leaves == 3


def test_weighted_counting_leaves_are_rejected_by_plans_and_exact_step_by_step():
    """A plan stages its leaves as 0/1 bytes and sizes every intermediate from that bound: counting boxes with
    other entries must not be truncated silently.  SimpleContraction (one contract2 per step, int64) is exact."""
    import torch  # noqa: F401

    from tensorsat_b200 import _native
    from tensorsat_b200.diagrams import Diagram, Wiring

    wiring = Wiring(wire_types=[FinSet(2)] * 3, slot_wires_list=[[0, 1], [1, 2]], out_wires=[0, 2])
    a = np.array([[1, 2], [3, 300]], dtype=np.int64)
    b = np.array([[5, 1], [4, 5]], dtype=np.int64)
    boxes = (CudaCountRel(a), CudaCountRel(b))
    assert not boxes[0].is_binary() and CudaCountRel(np.ones((2, 2), np.int64)).is_binary()
    diagram = Diagram(wiring, boxes)
    want = a @ b
    got = SimpleContraction(CudaCountRel, wiring, [(0, 1)]).contract(diagram)
    assert np.array_equal(got.tensor, want) and int(got.tensor[1, 1]) == 1503
    with pytest.raises(ValueError, match="0/1"):
        TreeContraction(CudaCountRel, wiring, ssa_path=[(0, 1)]).contract(diagram)
    # a device-produced count tensor fed back as a leaf is checked on the device (tsb_all_binary)
    fed = (got, CudaCountRel(np.ones((2, 2), np.int64)))
    assert got._binary is None
    with pytest.raises(ValueError, match="0/1"):
        TreeContraction(CudaCountRel, wiring, ssa_path=[(0, 1)]).contract(Diagram(wiring, fed))
    assert got._binary is False
    ok = CudaCountRel.contract2(CudaCountRel(np.eye(2, dtype=np.int64)), [0, 1], CudaCountRel(np.eye(2, dtype=np.int64)), [1, 2], [0, 2])
    assert ok._binary is None and ok.is_binary()
    res = TreeContraction(CudaCountRel, wiring, ssa_path=[(0, 1)]).contract(Diagram(wiring, (ok, CudaCountRel(np.ones((2, 2), np.int64)))))
    assert np.array_equal(res.tensor, np.ones((2, 2), np.int64))
    assert _native.all_binary(_native.upload(np.array([0, 1, 1, 0], np.uint8)))
    assert not _native.all_binary(_native.upload(np.array([0, 1, -1, 0], np.int64)))


def test_witness_pinned_subcubes_are_exact():
    """Sampled sub-cubes (planner.sampling / planner.subcube): pins that follow a known model give a sub-cube whose
    answer is SAT with the exact sub-cube count; single flips and random pins agree with brute force and with the
    independent exact counter (oracle/sharp_sat.py) — the small-scale twin of tests/test_gpu_wide_configs.py."""
    from oracle.sharp_sat import count_models, network_to_cnf
    from tensorsat_b200 import _native
    from tensorsat_b200.planner import propagate_wire_values, sample_subcube

    n = 16
    cnf = CNFInstance.random(3, n, 64, rng=9)
    brute = O.cnf_brute_force(n, cnf.clauses).reshape((2,) * n)
    model = [int(x) for x in np.argwhere(brute)[0]]
    flat, net = _network_and_leaves(cnf.inputs(None) >> cnf.diagram())
    leaves = [b.tensor for b in flat.boxes]
    vals = propagate_wire_values(net, leaves, dict(enumerate(model)))
    assert len(vals) == net.num_wires
    tree = find_tree_elimination(net, trials=2, seed=0).slice_to(6)
    assert len(tree.sliced_wires) >= 4
    wcnf = network_to_cnf(net.inputs, leaves)
    free = [w < n for w in range(net.num_wires)]
    for witness, seed in ((vals, 0), (None, 1), (None, 2)):
        sub, subleaves, asg = sample_subcube(tree, leaves, 2, seed=seed, witness=witness)
        cases = [asg] + [{**asg, w: asg[w] ^ 1} for w in list(asg)[:3]]
        for a in cases:
            from tensorsat_b200.planner.sampling import fix_wires

            net_a, leaves_a = fix_wires(net, leaves, a)
            want = count_models(net.num_wires, wcnf, a)
            if all(w < n for w in a):
                idx = tuple(a.get(v, slice(None)) for v in range(n))
                assert want == int(brute[idx].sum()) == count_models(n, cnf.clauses, a)
            for mode in (_native.MODE_SAT, _native.MODE_COUNT):
                plan = _native.Plan(net_a.inputs, net_a.output, net_a.size_dict, tree.ssa_path, sub.sliced_wires, mode=mode,
                                    wire_is_free=free)
                out = _native.empty(1, mode)
                plan.execute([_native.upload(np.ascontiguousarray(t)) for t in leaves_a], out)
                got = int(out.cpu().numpy()[0])
                assert got == (want if mode == _native.MODE_COUNT else int(want > 0)), (a, mode, got, want)
        if witness is not None:
            assert count_models(net.num_wires, wcnf, asg) >= 1


def test_slice_blocks_restrict_the_once_only_tensors():
    """A rank that executes an aligned power-of-two block of slice ids computes only its own sub-blocks of the
    once-only tensors that keep sliced wires (tsb_plan.cu: pre_block).  Plans large enough for the sub-block
    launches to happen, both semirings: every block alone, and the blocks of a partition combined (OR / sum),
    must give what the full range gives, and that must be brute force."""
    import torch

    from tensorsat_b200 import _native

    n, m, seed = 22, 86, 11
    cnf = CNFInstance.random(3, n, m, rng=seed)
    flat, net0 = _network_and_leaves(cnf.inputs(None) >> cnf.diagram())
    outs = (0, 7, 15, 21)
    net = Network(net0.inputs, outs, net0.size_dict)
    tree = find_tree_elimination(net, trials=2, seed=seed)
    tree = tree.slice_to(tree.width() - 3)
    assert tree.num_slices == 8
    brute = O.cnf_brute_force(n, cnf.clauses).reshape((2,) * n)
    rest = tuple(v for v in range(n) if v not in outs)
    leaves = [_native.upload(b.tensor.astype(np.uint8)) for b in flat.boxes]
    free = [w < n for w in range(net.num_wires)]
    for mode, want in ((_native.MODE_SAT, brute.any(axis=rest).astype(np.uint8)), (_native.MODE_COUNT, brute.sum(axis=rest, dtype=np.int64))):
        plan = _native.Plan(net.inputs, outs, net.size_dict, tree.ssa_path, tree.sliced_wires, mode=mode, batch_log2=30,
                            wire_is_free=free if mode == _native.MODE_COUNT else None)
        full = _native.empty(plan.info["out_numel"], mode)
        plan.execute(leaves, full)
        assert plan.describe()["restricted_launches"] == 0
        assert np.array_equal(full.cpu().numpy().reshape(want.shape), want)
        for size in (4, 2, 1):
            parts = []
            for lo in range(0, 8, size):
                r = torch.zeros_like(full)
                plan.execute(leaves, r, lo, lo + size)
                parts.append(r.cpu().numpy().reshape(want.shape).copy())
            assert plan.describe()["restricted_launches"] > 0   # packed steps / conversions (SAT), typed merges (counting)
            comb = np.bitwise_or.reduce(parts) if mode == _native.MODE_SAT else np.sum(parts, axis=0)
            assert np.array_equal(comb, want), (mode, size)
        # an unaligned range is executed without the restriction
        r = torch.zeros_like(full)
        plan.execute(leaves, r, 1, 3)
        assert plan.describe()["restricted_launches"] == 0


def test_first_model_reproduces_the_reference_traces():
    """``AssignmentSearch`` (one plan, leaves swapped per decision) against the reference notebooks: the sat / unsat
    trace of 01-introduction.ipynb:1848-1849 (G5), the lexicographically first model of 02-random-ksat.ipynb:707 (G4,
    on the GPU only: 30 executions of a width-18 plan are slow on the emulation), brute force otherwise."""
    from tensorsat_b200 import _native
    from tensorsat_b200.lib.search import AssignmentSearch

    cnf = CNFInstance.random(3, 3, 6, rng=3)
    for propagate in (False, True):   # pure: every decision is a contraction; default: unit propagation answers what it can
        search = AssignmentSearch(cnf, TreeContraction(CudaFinRel, cnf.contraction_wiring()), propagate=propagate)
        trace = []
        assert search.first_model(lambda spec, ok: trace.append((spec, ok))) == "011"
        assert trace == [("???", True), ("0??", True), ("00?", False), ("01?", True), ("010", False), ("011", True)]
        assert search.decisions + search.implied == 6 and (search.implied > 0) == propagate
    if _native._device == "cuda":
        cnf = CNFInstance.random(4, 20, 195, rng=1)
        search = AssignmentSearch(cnf, TreeContraction(CudaFinRel, cnf.contraction_wiring(), trials=2), propagate=False)
        trace = []
        assert search.first_model(lambda spec, ok: trace.append((spec, ok))) == "10011001000011100101" and search.decisions == 30
        fast = AssignmentSearch(cnf, TreeContraction(CudaFinRel, cnf.contraction_wiring(), trials=2))
        trace2 = []
        assert fast.first_model(lambda spec, ok: trace2.append((spec, ok))) == "10011001000011100101"
        assert trace2 == trace and fast.decisions + fast.implied == 30 and fast.decisions <= 10
    n = 11
    cnf = CNFInstance.random(3, n, 44, rng=2)
    brute = O.cnf_brute_force(n, cnf.clauses).reshape(-1)
    want = format(int(np.flatnonzero(brute)[0]), "b").rjust(n, "0") if brute.any() else None
    traces = []
    for propagate in (False, True):
        search = AssignmentSearch(cnf, TreeContraction(CudaFinRel, cnf.contraction_wiring(), trials=2), propagate=propagate)
        traces.append([])
        assert search.first_model(lambda spec, ok: traces[-1].append((spec, ok))) == want
    assert traces[0] == traces[1]
    import os

    from tests.conftest import GOLDEN

    with open(os.path.join(GOLDEN, "kissat", "full2.cnf")) as f:
        unsat = CNFInstance.from_dimacs(f.read())
    assert AssignmentSearch(unsat, TreeContraction(CudaFinRel, unsat.contraction_wiring())).first_model() is None
    with pytest.raises(ValueError):
        search.satisfiable("0" * (n - 1))
