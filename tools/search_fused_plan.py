#!/usr/bin/env python
"""
Tree search by the plan compiler's own cost: randomised elimination orders are
sliced and COMPILED (tsb_plan_create on the host-emulated build — plan construction needs no GPU), and ranked by a
model of what the executor is bound by on B200:

    time ~ sum over per-slice launches of (operand fetches + 1 store) x 16 B x output vectors x n_slices / L1_RATE
           + once-only launches x T_ONCE + per-slice launches x n_slices x T_SLICE_LAUNCH

with L1_RATE = 14.1 TB/s, T_ONCE = 3.6 us, T_SLICE_LAUNCH = 1 us, calibrated on the committed n=50 bench plan
(slice loop 1.053 ms for 14.9 GB of L1 traffic, once-only part 0.30 ms for 83 launches; profiles/README.md).
Step-level flops / bytes objectives rank these trees badly because fusion changes what is stored.

    python tools/search_fused_plan.py --trials 400 --out bench_plans/candidates.json
"""
import argparse
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

L1_RATE, T_ONCE, T_SLICE_LAUNCH = 14.1e12, 3.6e-6, 1.0e-6


def model_cost(plan):
    d = plan.describe()
    n_slices = plan.info["n_slices"]
    l1 = 0.0
    n_slice_launch = 0
    for o in d["ops"]:
        if not o["per_slice"]:
            continue
        n_slice_launch += 1
        vec = 2.0 ** max(o["out_bits"] - 7, 0)
        if o["kind"] == "bits":
            fetch = o["fetches_x4"] / 4.0 if o["fetches_x4"] else 2.0 * (2 ** o["k"])
            l1 += (fetch + 1.0) * 16.0 * vec
        else:   # byte kernels and conversions: the bytes they move
            l1 += o["bytes"]
    t = l1 * n_slices / L1_RATE + plan.info["n_kernels_shared"] * T_ONCE + n_slice_launch * n_slices * T_SLICE_LAUNCH
    return t, {"l1_bytes_per_slice": l1, "n_slices": int(n_slices), "launches_per_slice": n_slice_launch,
               "launches_once": int(plan.info["n_kernels_shared"]), "peak_bytes": int(plan.info["peak_bytes"])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=200)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--targets", default="30,31,32")
    ap.add_argument("--out", default=os.path.join(ROOT, "bench_plans", "candidates.json"))
    args = ap.parse_args()
    from tests.backend import bind_emulation

    bind_emulation()
    import bench
    from tensorsat_b200 import _native
    from tensorsat_b200.planner import ContractionTree
    from tensorsat_b200.planner.search import _elimination_order, _tree_from_order

    cnf, diagram, base = bench.load_tree()
    net = base.network

    def evaluate(tree):
        plan = _native.Plan(net.inputs, net.output, net.size_dict, tree.ssa_path, tree.sliced_wires, mode=_native.MODE_SAT)
        try:
            return model_cost(plan)
        finally:
            plan.close()

    t0, meta0 = evaluate(base)
    print(f"committed bench plan: model {t0 * 1e3:.3f} ms {meta0}", flush=True)
    best = []
    rng = random.Random(args.seed)
    targets = [int(x) for x in args.targets.split(",")]
    start = time.time()
    for t in range(args.trials):
        heuristic = "min_fill" if t % 2 == 0 else "min_degree"
        noise = 0.0 if t < 2 else rng.choice([0.01, 0.02, 0.05, 0.1, 0.2, 0.4])
        order = _elimination_order(net, rng, heuristic, noise)
        path = _tree_from_order(net, order)
        tree0 = ContractionTree(net, path)
        if tree0.width() > 38:
            continue
        for target in targets:
            tree = tree0.slice_to_fast(target) if tree0.width() > target else tree0
            try:
                cost, meta = evaluate(tree)
            except Exception as e:   # a tree the plan compiler refuses
                print("skip", e, flush=True)
                continue
            if meta["peak_bytes"] > 24 << 30:
                continue
            best.append((cost, t, target, tree, meta))
        best.sort(key=lambda x: x[0])
        best = best[:5]
        if t % 20 == 0:
            print(f"trial {t} ({time.time() - start:.0f} s): best model {best[0][0] * 1e3:.3f} ms target {best[0][2]} {best[0][4]}", flush=True)
    out = [{"model_ms": c * 1e3, "trial": t, "target": target, "meta": meta, "tree": tree.to_json()} for c, t, target, tree, meta in best]
    with open(args.out, "w") as f:
        json.dump({"committed_model_ms": t0 * 1e3, "candidates": out}, f, separators=(",", ":"))
    print("written", args.out, [round(o["model_ms"], 3) for o in out])


if __name__ == "__main__":
    main()
