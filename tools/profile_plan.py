#!/usr/bin/env python
"""Per-kernel timing of the bench plan (CUDA events around every launch; tsb_plan_set_timing).

    python tools/profile_plan.py [--slices S] [--top T] [--batch-log2 B]

Prints the top ops by time with their algorithmic bytes and GB/s, and totals per kernel kind.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slices", type=int, default=0, help="number of slices to run (0 = all)")
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--batch-log2", type=int, default=-1)
    ap.add_argument("--json", default="")
    ap.add_argument("--count", action="store_true", help="counting semiring (#SAT) instead of SAT")
    args = ap.parse_args()
    import torch

    import bench
    from tensorsat_b200 import _native
    from tensorsat_b200.contractions.tree import TreeContraction
    from tensorsat_b200.lang.fin_rel import CudaCountRel, CudaFinRel

    cnf, diagram, tree = bench.load_tree()
    if args.count:
        c = TreeContraction(CudaCountRel, diagram.wiring, tree=tree, batch_log2=args.batch_log2,
                            wire_is_free=[w < cnf.num_vars for w in diagram.wiring.wires])
    else:
        c = TreeContraction(CudaFinRel, diagram.wiring, tree=tree, batch_log2=args.batch_log2)
    plan = c.plan()
    hi = plan.info["n_slices"] if args.slices <= 0 else min(args.slices, plan.info["n_slices"])
    leaves = [c._leaf_tensor(b) for b in diagram.boxes]
    out = _native.empty(plan.info["out_numel"], c.mode)
    for _ in range(2):
        plan.execute(leaves, out, 0, hi)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan.execute(leaves, out, 0, hi)
    e1.record()
    torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1)
    plan.set_timing(True)
    plan.execute(leaves, out, 0, hi)
    torch.cuda.synchronize()
    d = plan.describe()
    ops = d["ops"]
    kinds = {}
    for o in ops:
        mult = hi if o["per_slice"] else 1
        k = kinds.setdefault((o["kind"], o["per_slice"]), {"n": 0, "ms": 0.0, "bytes": 0.0})
        k["n"] += mult
        k["ms"] += o["ms"]
        k["bytes"] += o["bytes"] * mult
    print(f"untimed execution: {total_ms:.3f} ms for {hi} slice(s); result={out.cpu().numpy().reshape(-1)[:4]} info={plan.info}")
    for (kind, ps), k in sorted(kinds.items()):
        gbps = k["bytes"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] else 0
        print(f"{kind:8s} per_slice={ps} launches={k['n']:6d} ms={k['ms']:9.3f} alg_GB={k['bytes']/1e9:9.3f} GB/s={gbps:8.1f}")
    ops_sorted = sorted(ops, key=lambda o: -o["ms"])[: args.top]
    for o in ops_sorted:
        mult = hi if o["per_slice"] else 1
        gbps = o["bytes"] * mult / (o["ms"] * 1e-3) / 1e9 if o["ms"] else 0
        print(f"{o['kind']:8s} ps={o['per_slice']} dt={o['dtype']} out_bits={o['out_bits']:2d} b={o['b']:2d} m={o['m']:2d} n={o['n']:2d} k={o['k']} "
              f"ms(total over slices)={o['ms']:8.3f} per-launch us={o['ms']/mult*1e3:9.1f} GB/s={gbps:8.1f}")
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"total_ms": total_ms, "slices": hi, "ops": ops, "info": plan.info}, f)


if __name__ == "__main__":
    main()
