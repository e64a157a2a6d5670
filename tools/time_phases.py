#!/usr/bin/env python
"""Device time of the bench plan for slice ranges [0, k): k = 0 is the once-only part alone (graph replay, as in a run).

    python tools/time_phases.py [--reps 30]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=30)
    args = ap.parse_args()
    import torch

    import bench
    from tensorsat_b200 import _native
    from tensorsat_b200.contractions.tree import TreeContraction
    from tensorsat_b200.lang.fin_rel import CudaFinRel

    cnf, diagram, tree = bench.load_tree()
    c = TreeContraction(CudaFinRel, diagram.wiring, tree=tree)
    plan = c.plan()
    leaves = [c._leaf_tensor(b) for b in diagram.boxes]
    out = _native.empty(plan.info["out_numel"], c.mode)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for hi in (0, 1, 2, 16):
        for _ in range(4):
            plan.execute(leaves, out, 0, hi)
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.execute(leaves, out, 0, hi)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        print(f"slices [0, {hi}): median {ts[len(ts) // 2]:.4f} ms  min {ts[0]:.4f}  graph {plan.describe().get('graph_replays')}", flush=True)


if __name__ == "__main__":
    main()
