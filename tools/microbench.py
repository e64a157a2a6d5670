#!/usr/bin/env python
"""Kernel micro-benchmarks: synthetic 0/1 tensors through the C-ABI, timed with CUDA
events on the launching stream after warm-up; every tensor is larger than L2.

    python tools/microbench.py merge   [--bits 30]     absorb / merge steps shaped like the C1 plan's top steps
    python tools/microbench.py permute [--bits 30]     bit-permutation transposes (plan-like and random)
    python tools/microbench.py gemm                    tcgen05 int8 GEMM sweep
Each line is a JSON record with algorithmic GB/s (or TOP/s) and the fraction of the measured peak.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), 2.0 * float(d["bf16_tflops"]), "measured"
    return 6650.0, 2 * 1590.0, "fallback"


def time_op(fn, iters=10, warmup=3):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def rand01(n, p=0.5, seed=0):
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.rand(n, device="cuda", generator=g) < p).to(torch.uint8)


def bench_merge(args):
    import torch

    from tensorsat_b200 import _native

    hbm, _, src = peaks()
    R = args.bits
    rng = np.random.default_rng(0)
    cases = []
    # (name, big wires, small wires, out wires) in one global order: wire id = significance rank (0 = msb)
    # absorb step, k = 1: big [k, R wires], small [k, b batch wires, n new wires]
    for nb, nn in ((21, 12), (26, 6), (10, 22), (8, 2), (12, 0)):
        nb = min(nb, R - nn) if nn else min(nb, R)
        out = list(range(1, R + 1))
        new = sorted(int(x) for x in rng.choice(out, size=nn, replace=False))
        big = [0] + [w for w in out if w not in new]
        batch = sorted(int(x) for x in rng.choice([w for w in out if w not in new], size=min(nb, R - nn), replace=False))
        small = [0] + sorted(batch + new)
        cases.append((f"absorb b={len(batch)} n={nn} k=1", big, small, out))
    # write-dominated outer-product-like step (b=17 m=8 n=9 k=1 in the C1 plan)
    out = list(range(1, R + 1))
    perm = [int(x) for x in rng.permutation(out)]
    nb_, nm_ = R // 2, R // 4
    b, m, n = perm[:nb_], perm[nb_:nb_ + nm_], perm[nb_ + nm_:]
    cases.append((f"outer b={len(b)} m={len(m)} n={len(n)} k=1", [0] + sorted(b + m), [0] + sorted(b + n), out))
    if args.from_plan:
        # replay the largest merge steps of the bench plan with synthetic data of the same label structure
        import bench
        from tensorsat_b200.contractions.tree import TreeContraction
        from tensorsat_b200.lang.fin_rel import CudaFinRel

        cnf, diagram, tree = bench.load_tree()
        desc = TreeContraction(CudaFinRel, diagram.wiring, tree=tree).plan().describe()
        ops = sorted([o for o in desc["ops"] if "ow" in o and o["kind"] == "merge"], key=lambda o: -o["bytes"])[: args.from_plan]
        cases = [(f"plan#{i} out={len(o['ow'])} b={o['b']} m={o['m']} n={o['n']} k={o['k']} nibL={o['nibL']} nibR={o['nibR']}",
                  o["lw"], o["rw"], o["ow"]) for i, o in enumerate(ops)]
    if args.case:
        cases = [c for c in cases if args.case in c[0]]
    for name, lw, rw, ow in cases:
        L = rand01(1 << len(lw), 0.4, 1)
        Rr = rand01(1 << len(rw), 0.6, 2)
        O = torch.empty(1 << len(ow), dtype=torch.uint8, device="cuda")
        fn = lambda: _native.contract2(L, [2] * len(lw), lw, Rr, [2] * len(rw), rw, O, [2] * len(ow), ow, 0)  # noqa: E731
        ms = time_op(fn, iters=args.iters)
        nbytes = L.numel() + Rr.numel() + O.numel()
        gbps = nbytes / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": "k_merge<u8,u8,u8,clamp>", "case": name, "lhs_bits": len(lw), "rhs_bits": len(rw),
                          "out_bits": len(ow), "ms": ms, "alg_bytes": nbytes, "GBps": gbps, "frac_hbm": gbps / hbm,
                          "peak": hbm, "peak_source": src}))
        del L, Rr, O


def bench_permute(args):
    import torch

    from tensorsat_b200 import _native

    hbm, _, src = peaks()
    r = args.bits
    rng = np.random.default_rng(0)
    src_t = rand01(1 << r, 0.5, 3)
    dst_t = torch.empty_like(src_t)
    perms = {"reverse": list(range(r))[::-1], "rotate_by_7": list(range(7, r)) + list(range(7)),
             "swap_halves": list(range(r // 2, r)) + list(range(r // 2)),
             "keep_low_4": [int(x) for x in rng.permutation(r - 4)] + list(range(r - 4, r))}
    for i in range(args.random):
        perms[f"random_{i}"] = [int(x) for x in rng.permutation(r)]
    for name, perm in perms.items():
        fn = lambda: _native.permute(src_t, dst_t, [2] * r, perm, 1)  # noqa: E731
        ms = time_op(fn, iters=args.iters)
        gbps = 2.0 * (1 << r) / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": "k_permute", "case": name, "bits": r, "elem": 1, "ms": ms, "alg_bytes": 2 << r,
                          "GBps": gbps, "frac_hbm": gbps / hbm, "peak": hbm, "peak_source": src}))
    # 8-byte elements (counting mode)
    r8 = r - 3
    s8 = torch.arange(1 << r8, dtype=torch.int64, device="cuda")
    d8 = torch.empty_like(s8)
    for name, perm in {"i64_reverse": list(range(r8))[::-1], "i64_random": [int(x) for x in rng.permutation(r8)]}.items():
        fn = lambda: _native.permute(s8, d8, [2] * r8, perm, 8)  # noqa: E731
        ms = time_op(fn, iters=args.iters)
        gbps = 2.0 * 8 * (1 << r8) / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": "k_permute", "case": name, "bits": r8, "elem": 8, "ms": ms, "GBps": gbps,
                          "frac_hbm": gbps / hbm, "peak": hbm, "peak_source": src}))
    # copy / write-only / read-only ceilings on this box, for context
    ms = time_op(lambda: dst_t.copy_(src_t), iters=args.iters)
    print(json.dumps({"kernel": "torch copy_ (context)", "bits": r, "ms": ms, "GBps": 2.0 * (1 << r) / (ms * 1e-3) / 1e9}))
    ms = time_op(lambda: dst_t.fill_(1), iters=args.iters)
    print(json.dumps({"kernel": "torch fill_ write-only (context)", "bits": r, "ms": ms, "GBps": (1 << r) / (ms * 1e-3) / 1e9}))
    s32 = src_t.view(torch.int32)
    ms = time_op(lambda: s32.sum(), iters=args.iters)
    print(json.dumps({"kernel": "torch sum read-only (context)", "bits": r, "ms": ms, "GBps": (1 << r) / (ms * 1e-3) / 1e9}))


def bench_gemm(args):
    import torch

    from tensorsat_b200 import _native

    _, tops_peak, src = peaks()
    for (b, m, n, k) in [(0, 13, 13, 13), (0, 12, 12, 12), (4, 11, 11, 11), (8, 10, 10, 9), (10, 9, 9, 8), (12, 8, 8, 7),
                         (0, 14, 14, 10), (0, 15, 15, 6)]:
        A = rand01(1 << (b + m + k), 0.5, 1)
        B = rand01(1 << (b + n + k), 0.5, 2)
        for mode, dt in ((0, torch.uint8), (1, torch.int32)):
            Cc = torch.empty(1 << (b + m + n), dtype=dt, device="cuda")
            fn = lambda: _native.gemm_u8(A, B, Cc, b, m, n, k, mode)  # noqa: E731
            ms = time_op(fn, iters=args.iters)
            ops = 2.0 * 2.0 ** (b + m + n + k)
            nbytes = A.numel() + B.numel() + Cc.numel() * Cc.element_size()
            print(json.dumps({"kernel": "k_gemm_i8_tcgen05", "b": b, "m": m, "n": n, "k": k, "mode": "sat" if mode == 0 else "count",
                              "ms": ms, "TOPs": ops / (ms * 1e-3) / 1e12, "frac_int8_peak": ops / (ms * 1e-3) / 1e12 / tops_peak,
                              "peak_TOPs": tops_peak, "peak_source": src + " (2 x bf16 dense)", "GBps": nbytes / (ms * 1e-3) / 1e9,
                              "intensity_op_per_B": ops / nbytes}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["merge", "permute", "gemm"])
    ap.add_argument("--bits", type=int, default=30)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--random", type=int, default=4)
    ap.add_argument("--case", default="")
    ap.add_argument("--from-plan", type=int, default=0)
    a = ap.parse_args()
    {"merge": bench_merge, "permute": bench_permute, "gemm": bench_gemm}[a.what](a)
