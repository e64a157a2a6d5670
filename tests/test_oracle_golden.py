"""Pins the numpy oracle against vectors recorded from the reference's own code
(tests/golden/make_golden.py) and against the known answers of SURVEY.md §8c."""

import numpy as np
import pytest

from oracle import fin_rel_oracle as O


def _arr(x, shape=None, dtype=np.uint8):
    a = np.array(x, dtype=dtype)
    return a.reshape(shape) if shape is not None else a


def _surviving_in_order(ports):
    from collections import Counter

    c = Counter(ports)
    single = [p for p in ports if c[p] == 1]
    return single == sorted(single)


def test_contract2_matches_reference_verbatim(golden_ops):
    n = 0
    for c in golden_ops:
        if c["op"] != "contract2":
            continue
        lhs, rhs = _arr(c["lhs"], c["lhs_shape"]), _arr(c["rhs"], c["rhs_shape"])
        ow = c["out_wires"]
        if ow is None:
            ow = sorted(set(c["lhs_wires"]).symmetric_difference(c["rhs_wires"]))
        if "raises" in c:
            with pytest.raises(ValueError):
                O.contract2(lhs, c["lhs_wires"], rhs, c["rhs_wires"], ow,
                            ref_quirk_mod256=True, ref_quirk_rewire=True)
            continue
        got = O.contract2(lhs, c["lhs_wires"], rhs, c["rhs_wires"], ow,
                          ref_quirk_mod256=True, ref_quirk_rewire=True)
        want = _arr(c["res"], c["res_shape"])
        assert got.dtype == np.uint8 and got.shape == want.shape
        assert np.array_equal(got, want)
        n += 1
    assert n >= 40


def test_contract2_exact_semantics_agree_where_reference_is_well_defined(golden_ops):
    n = 0
    for c in golden_ops:
        if c["op"] != "contract2" or "raises" in c:
            continue
        ow = c["out_wires"]
        if ow is None:
            ow = sorted(set(c["lhs_wires"]).symmetric_difference(c["rhs_wires"]))
        if len(set(ow)) != len(ow):
            rp = [sorted(set(ow)).index(w) for w in ow]
            if not _surviving_in_order(rp):
                continue  # reference quirk R2 territory
        lhs, rhs = _arr(c["lhs"], c["lhs_shape"]), _arr(c["rhs"], c["rhs_shape"])
        got = O.contract2(lhs, c["lhs_wires"], rhs, c["rhs_wires"], ow)
        assert np.array_equal(got, _arr(c["res"], c["res_shape"]))
        n += 1
    assert n >= 35


def test_mod256_quirk_is_reproduced_and_explained(golden_ops):
    q = {c["k"]: c for c in golden_ops if c["op"] == "contract2_quirk"}
    ones8 = np.ones((2,) * 8, dtype=np.uint8)
    w = list(range(8))
    assert q[8]["res"] == 0  # the reference says "unsatisfiable" for 256 witnesses
    assert int(O.contract2(ones8, w, ones8, w, [], ref_quirk_mod256=True)) == 0
    assert int(O.contract2(ones8, w, ones8, w, [])) == 1
    assert int(O.contract2(ones8, w, ones8, w, [], mode="count")) == 256
    ones9 = np.ones((2,) * 9, dtype=np.uint8)
    w9 = list(range(9))
    assert np.array_equal(O.contract2(ones9, w9, ones9, w9, [8], ref_quirk_mod256=True), _arr(q[9]["res"]))
    assert np.array_equal(O.contract2(ones9, w9, ones9, w9, [8]), np.ones(2, dtype=np.uint8))


def test_rewire_matches_reference(golden_ops):
    n = 0
    for c in golden_ops:
        if c["op"] != "rewire":
            continue
        t = _arr(c["tensor"], c["shape"])
        if "raises" in c:
            with pytest.raises(ValueError):
                O.rewire(t, c["ports"], ref_quirk_mod256=True, ref_quirk_rewire=True)
            continue
        want = _arr(c["res"], c["res_shape"])
        got = O.rewire(t, c["ports"], ref_quirk_mod256=True, ref_quirk_rewire=True)
        assert np.array_equal(got, want)
        if _surviving_in_order(c["ports"]) or len(set(c["ports"])) == len(c["ports"]) == t.ndim:
            assert np.array_equal(O.rewire(t, c["ports"]), want)
        n += 1
    assert n >= 25


def test_small_ops_match_reference(golden_ops):
    for c in golden_ops:
        if c["op"] == "spider":
            assert np.array_equal(O.spider(c["size"], c["num_ports"]), _arr(c["res"]))
        elif c["op"] == "scalar":
            assert np.array_equal(O.scalar(c["value"]), _arr(c["res"]))
            assert O.as_scalar(O.scalar(c["value"])) == c["bool"]
        elif c["op"] == "slice_union_stack":
            a, b = _arr(c["a"], c["shape"]), _arr(c["b"], c["shape"])
            idx = tuple(slice(None) if i is None else i for i in c["idx"])
            assert np.array_equal(O.getitem(a, idx), _arr(c["getitem"], c["getitem_shape"]))
            assert np.array_equal(O.union(a, b), _arr(c["union"]))
            assert np.array_equal(O.stack([a, b, a], c["port"]), _arr(c["stack"]))
            assert np.array_equal(O.stack([a, b], c["port"]), _arr(c["autoray_stack"]))
        elif c["op"] == "einsum":
            lw, rw, ow = O.parse_einsum(c["eq"])
            got = O.contract2(_arr(c["lhs"]), lw, _arr(c["rhs"]), rw, ow)
            assert np.array_equal(got, _arr(c["res"]))
        elif c["op"] == "transpose":
            assert np.array_equal(O.rewire(_arr(c["tensor"]), c["perm"]), _arr(c["res"]))


def test_notebook_known_answers():
    """SURVEY.md §8c G6-G8, G12 (notebooks/01-introduction.ipynb, 2025-02-13-#sat prototype)."""
    and_ = _arr([[[1, 0], [1, 0]], [[1, 0], [0, 1]]])  # G6
    or_ = _arr([[[1, 0], [0, 1]], [[0, 1], [0, 1]]])
    xor_ = _arr([[[1, 0], [0, 1]], [[0, 1], [1, 0]]])
    not_ = _arr([[0, 1], [1, 0]])
    true = _arr([0, 1])
    pts = lambda t: {tuple(map(int, i)) for i in np.argwhere(t)}  # noqa: E731
    assert pts(O.rewire(and_, [0, 2, 1])) == {(0, 0, 0), (0, 0, 1), (1, 0, 0), (1, 1, 1)}  # G7
    assert np.array_equal(O.rewire(xor_, [0, 2, 1]), xor_)
    cnot = O.rewire(xor_, [0, 1, 0, 2])
    assert pts(cnot) == {(a, b, a, a ^ b) for a in (0, 1) for b in (0, 1)}
    assert np.array_equal(O.rewire(cnot, [0, 1, 3]), xor_)
    sp3 = O.spider(2, 3)  # G8
    assert np.array_equal(O.contract2(sp3, [0, 2, 4], xor_, [4, 1, 3], [0, 1, 2, 3]), cnot)
    got = O.contract2(or_, [1, 2, 3], and_, [0, 3, 4], [0, 1, 2, 4])
    assert pts(got) == {(a, b, c, a & (b | c)) for a in (0, 1) for b in (0, 1) for c in (0, 1)}
    # G12: #SAT prototype, one einsum over the whole network, no clamp
    ops = [or_, or_, true, not_, not_, not_, or_, or_, true]
    assert int(np.einsum("abd,dce,e,cf,bg,ah,fgi,ihj,j->", *ops)) == 6
    assert np.einsum("abd,dce,e,cf,bg,ah,fgi,ihj,j->abc", *ops).reshape(-1).tolist() == [0, 1, 1, 1, 1, 1, 1, 0]
    inputs = [(0, 1, 3), (3, 2, 4), (4,), (2, 5), (1, 6), (0, 7), (5, 6, 8), (8, 7, 9), (9,)]
    path = [(0, 1), (9, 2), (10, 3), (11, 4), (12, 5), (13, 6), (14, 7), (15, 8)]
    assert int(O.execute_tree(inputs, (), ops, path, mode="count")) == 6
    assert O.execute_tree(inputs, (0, 1, 2), ops, path, mode="count").reshape(-1).tolist() == [0, 1, 1, 1, 1, 1, 1, 0]
    assert bool(O.as_scalar(O.execute_tree(inputs, (), ops, path)))
