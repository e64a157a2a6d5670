"""
Contraction-tree search: randomised greedy agglomeration with restarts.

Stand-in for cotengra's hyper-optimiser, which the reference calls through
``ctg.array_contract_tree(..., optimize=...)`` (tensorsat/contractions/cotengra.py:78-85;
``optimize="greedy"`` in examples/test_sat_solver.py:8).  Host-side, seedable,
deterministic for a given seed; not part of the timed contraction phase.

Hyper-edges are first class: a wire may touch any number of tensors and is summed only
when its last two holders meet (and it is not an output wire).
"""

from __future__ import annotations

import heapq
import math
import random
from typing import Sequence

from .network import Network
from .tree import ContractionTree, SSAPath


def _greedy_once(
    net: Network,
    rng: random.Random,
    alpha: float,
    temperature: float,
    sliced: frozenset[int] = frozenset(),
) -> SSAPath:
    L = net.num_leaves
    lg = [math.log2(d) for d in net.size_dict]
    wires_of: dict[int, frozenset[int]] = {
        i: frozenset(w for w in ws if w not in sliced) for i, ws in enumerate(net.inputs)
    }
    count: dict[int, int] = {}
    holders: dict[int, set[int]] = {}
    for i, ws in wires_of.items():
        for w in ws:
            count[w] = count.get(w, 0) + 1
            holders.setdefault(w, set()).add(i)
    for w in set(net.output) - sliced:
        count[w] = count.get(w, 0) + 1

    def size(ws: frozenset[int]) -> float:
        return 2.0 ** sum(lg[w] for w in ws)

    def out_of(a: int, b: int) -> frozenset[int]:
        la, lb = wires_of[a], wires_of[b]
        return frozenset(
            w for w in la | lb if count[w] - (1 if (w in la and w in lb) else 0) >= 2
        )

    def score(a: int, b: int) -> float:
        o = out_of(a, b)
        x = size(o) - alpha * (size(wires_of[a]) + size(wires_of[b]))
        s = math.copysign(math.log2(1.0 + abs(x)), x)
        if temperature > 0.0:
            u = rng.random()
            s -= temperature * -math.log(-math.log(u + 1e-300) + 1e-300)
        return s

    heap: list[tuple[float, int, int]] = []
    seen_pairs: set[tuple[int, int]] = set()

    def push_neighbours(a: int) -> None:
        nb: set[int] = set()
        for w in wires_of[a]:
            nb |= holders[w]
        nb.discard(a)
        for b in nb:
            p = (a, b) if a < b else (b, a)
            if p not in seen_pairs:
                seen_pairs.add(p)
                heapq.heappush(heap, (score(*p), p[0], p[1]))

    for i in range(L):
        push_neighbours(i)

    alive = set(range(L))
    path: list[tuple[int, int]] = []
    nxt = L
    while len(alive) > 1:
        pair = None
        while heap:
            _, a, b = heapq.heappop(heap)
            if a in alive and b in alive:
                pair = (a, b)
                break
        if pair is None:
            # disconnected components: combine the two smallest remaining tensors
            rest = sorted(alive, key=lambda t: (len(wires_of[t]), t))
            pair = (rest[0], rest[1])
        a, b = pair
        o = out_of(a, b)
        la, lb = wires_of[a], wires_of[b]
        for w in la | lb:
            holders[w].discard(a)
            holders[w].discard(b)
            if w in o:
                count[w] -= 1 if (w in la and w in lb) else 0
                holders[w].add(nxt)
            else:
                count[w] = 0
        alive.discard(a)
        alive.discard(b)
        alive.add(nxt)
        wires_of[nxt] = o
        path.append((a, b))
        push_neighbours(nxt)
        nxt += 1
    return tuple(path)


def _objective(tree: ContractionTree, minimize: str, max_log2_size: float | None) -> float:
    st = tree.stats()
    if minimize == "size":
        val = st.log2_max_size + 1e-3 * st.log2_ops
    elif minimize == "write":
        val = st.log2_write
    elif minimize == "flops":
        val = st.log2_ops
    else:  # "combo": ops + 64 * write, the usual memory-bound weighting
        val = math.log2(2.0**st.log2_ops + 64.0 * 2.0**st.log2_write)
    if max_log2_size is not None and st.log2_max_size > max_log2_size:
        val += 4.0 * (st.log2_max_size - max_log2_size)
    return val


def find_tree(
    net: Network,
    *,
    trials: int = 32,
    seed: int = 0,
    minimize: str = "combo",
    max_log2_size: float | None = None,
    alphas: Sequence[float] = (0.0, 0.5, 1.0, 1.5, 2.0),
    temperatures: Sequence[float] = (0.0, 0.05, 0.1, 0.2, 0.4, 0.8),
) -> ContractionTree:
    """Best of ``trials`` randomised greedy runs (trial 0 is the plain deterministic greedy)."""
    if net.num_leaves == 1:
        return ContractionTree(net, ())
    rng = random.Random(seed)
    best: tuple[float, ContractionTree] | None = None
    for t in range(max(1, trials)):
        if t == 0:
            alpha, temp = 1.0, 0.0
        else:
            alpha, temp = rng.choice(list(alphas)), rng.choice(list(temperatures))
        path = _greedy_once(net, rng, alpha, temp)
        tree = ContractionTree(net, path)
        obj = _objective(tree, minimize, max_log2_size)
        if best is None or obj < best[0]:
            best = (obj, tree)
    assert best is not None
    return best[1]


# --------------------------------------------------------------------------------------
# Elimination-order (bucket) trees
# --------------------------------------------------------------------------------------
def _elimination_order(
    net: Network, rng: random.Random, heuristic: str, noise: float
) -> list[int]:
    """Greedy vertex elimination on the wire primal graph (min-fill / min-degree)."""
    out = set(net.output)
    adj: dict[int, set[int]] = {w: set() for ws in net.inputs for w in ws}
    for ws in net.inputs:
        s = set(ws)
        for w in s:
            adj[w] |= s - {w}
    order: list[int] = []
    alive = set(adj)

    def fill(w: int) -> int:
        nb = list(adj[w])
        f = 0
        for i, a in enumerate(nb):
            aa = adj[a]
            for b in nb[i + 1 :]:
                if b not in aa:
                    f += 1
        return f

    cache: dict[int, float] = {}

    def key(w: int) -> float:
        if w not in cache:
            base = float(fill(w)) if heuristic == "min_fill" else float(len(adj[w]))
            cache[w] = base
        k = cache[w]
        if noise > 0.0:
            k += noise * rng.random() * (1.0 + k)
        return k

    while alive:
        cands = [w for w in alive if w not in out]
        if not cands:
            break
        w = min(cands, key=lambda x: (key(x), len(adj[x]), x))
        nb = adj[w]
        for a in nb:
            adj[a] |= nb - {a}
            adj[a].discard(w)
            cache.pop(a, None)
            for b in adj[a]:
                cache.pop(b, None)
        alive.discard(w)
        del adj[w]
        order.append(w)
    return order


def _tree_from_order(net: Network, order: Sequence[int]) -> SSAPath:
    """Bucket elimination: for each wire in order, merge every tensor that holds it."""
    L = net.num_leaves
    wires_of: dict[int, frozenset[int]] = {i: frozenset(ws) for i, ws in enumerate(net.inputs)}
    count: dict[int, int] = {}
    holders: dict[int, set[int]] = {}
    for i, ws in wires_of.items():
        for w in ws:
            count[w] = count.get(w, 0) + 1
            holders.setdefault(w, set()).add(i)
    for w in set(net.output):
        count[w] = count.get(w, 0) + 1
    path: list[tuple[int, int]] = []
    nxt = L
    alive = set(range(L))

    def merge(a: int, b: int) -> int:
        nonlocal nxt
        la, lb = wires_of[a], wires_of[b]
        o = set()
        for w in la | lb:
            c = count[w] - (1 if (w in la and w in lb) else 0)
            holders[w].discard(a)
            holders[w].discard(b)
            if c >= 2:
                o.add(w)
                count[w] = c
                holders[w].add(nxt)
            else:
                count[w] = 0
        wires_of[nxt] = frozenset(o)
        alive.discard(a)
        alive.discard(b)
        alive.add(nxt)
        path.append((a, b))
        nxt += 1
        return nxt - 1

    def out_rank(a: int, b: int) -> int:
        la, lb = wires_of[a], wires_of[b]
        return sum(1 for w in la | lb if count[w] - (1 if (w in la and w in lb) else 0) >= 2)

    for w in order:
        bucket = set(holders.get(w, ()))
        while len(bucket) > 1:
            # merge the pair with the smallest result (quadratic in the bucket, buckets are small)
            bl = sorted(bucket, key=lambda t: (len(wires_of[t]), t))
            best = None
            for i, a in enumerate(bl[:8]):
                for b in bl[i + 1 : 9]:
                    r = (out_rank(a, b), a, b)
                    if best is None or r < best:
                        best = r
            assert best is not None
            _, a, b = best
            bucket.discard(a)
            bucket.discard(b)
            bucket.add(merge(a, b))
    rest = sorted(alive, key=lambda t: (len(wires_of[t]), t))
    while len(rest) > 1:
        a, b = rest[0], rest[1]
        c = merge(a, b)
        rest = sorted(alive, key=lambda t: (len(wires_of[t]), t))
        del c
    return tuple(path)


def _anneal_order_tail(
    net: Network, order: Sequence[int], rng: random.Random, iters: int, tail: int = 64
) -> list[int]:
    """Simulated annealing on the LAST ``tail`` wires of an elimination order.

    Greedy min-fill / min-degree orders eliminate the cheap wires first (on a CNF network: the
    internal wires of the clause gadgets), and all of them end within a few percent of each
    other; what decides the cost is the order of the densely connected wires left at the end
    (the variables).  The head of the order is eliminated once on the wire primal graph; the
    tail is then permuted (swaps and single-wire moves) under the proxy cost
    sum over eliminated wires of 2^(weight of the wire and its neighbours at that moment),
    evaluated on bit sets.  Host-side, deterministic for a given seed.
    """
    order = list(order)
    T = min(tail, len(order))
    if T < 3 or iters <= 0:
        return order
    head, tl = order[: len(order) - T], order[len(order) - T :]
    adj: dict[int, set[int]] = {w: set() for ws in net.inputs for w in ws}
    for ws in net.inputs:
        s = set(ws)
        for w in s:
            adj[w] |= s - {w}
    for w in head:
        nb = adj.pop(w)
        for a in nb:
            adj[a] |= nb - {a}
            adj[a].discard(w)
    verts = list(tl) + sorted(w for w in adj if w not in set(tl))   # tail wires first, then output wires
    index = {w: i for i, w in enumerate(verts)}
    lg = [math.log2(net.size_dict[w]) for w in verts]
    uniform = all(abs(x - 1.0) < 1e-12 for x in lg)
    adj0 = [0] * len(verts)
    for w in verts:
        m = 0
        for a in adj[w]:
            m |= 1 << index[a]
        adj0[index[w]] = m
    full = (1 << len(verts)) - 1

    def weight(mask: int) -> float:
        if uniform:
            return float(bin(mask).count("1"))
        t = 0.0
        while mask:
            b = mask & -mask
            t += lg[b.bit_length() - 1]
            mask ^= b
        return t

    def cost(perm: Sequence[int]) -> float:
        a = adj0[:]
        alive = full
        tot = 0.0
        for v in perm:
            nb = a[v] & alive & ~(1 << v)
            tot += 2.0 ** (weight(nb) + lg[v])
            alive &= ~(1 << v)
            m = nb
            while m:
                b = m & -m
                a[b.bit_length() - 1] |= nb
                m ^= b
        return math.log2(tot)

    cur = list(range(T))
    cc = cost(cur)
    best, bc = cur[:], cc
    t_hi, t_lo = 0.3, 0.01
    for it in range(iters):
        temp = t_hi * (t_lo / t_hi) ** (it / iters)
        i, j = rng.randrange(T), rng.randrange(T)
        if i == j:
            continue
        new = cur[:]
        if rng.random() < 0.5:
            new[i], new[j] = new[j], new[i]
        else:
            new.insert(j, new.pop(i))
        c = cost(new)
        if c < cc or rng.random() < math.exp((cc - c) / temp):
            cur, cc = new, c
            if c < bc:
                best, bc = new[:], c
    return head + [tl[i] for i in best]


def find_tree_elimination(
    net: Network,
    *,
    trials: int = 16,
    seed: int = 0,
    minimize: str = "combo",
    max_log2_size: float | None = None,
    anneal_iters: int = 0,
    anneal_tail: int = 64,
    anneal_restarts: int = 1,
) -> ContractionTree:
    """Best of ``trials`` randomised min-fill / min-degree elimination orders; with ``anneal_iters`` > 0 the
    best order's tail is then refined by simulated annealing (kept only if the real objective improves)."""
    if net.num_leaves == 1:
        return ContractionTree(net, ())
    rng = random.Random(seed)
    best: tuple[float, ContractionTree] | None = None
    best_order: list[int] = []
    for t in range(max(1, trials)):
        heuristic = "min_fill" if t % 2 == 0 else "min_degree"
        noise = 0.0 if t < 2 else rng.choice([0.02, 0.05, 0.1, 0.2, 0.4])
        order = _elimination_order(net, rng, heuristic, noise)
        tree = ContractionTree(net, _tree_from_order(net, order))
        obj = _objective(tree, minimize, max_log2_size)
        if best is None or obj < best[0]:
            best = (obj, tree)
            best_order = order
    assert best is not None
    for _ in range(max(1, anneal_restarts) if anneal_iters > 0 else 0):
        # the proxy cost of the annealing tracks the real objective only loosely: every restart is judged by the
        # real one
        order = _anneal_order_tail(net, best_order, rng, anneal_iters, anneal_tail)
        tree = ContractionTree(net, _tree_from_order(net, order))
        obj = _objective(tree, minimize, max_log2_size)
        if obj < best[0]:
            best = (obj, tree)
    return best[1]
