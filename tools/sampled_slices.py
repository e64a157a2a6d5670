#!/usr/bin/env python
"""
Per-slice throughput of the BASELINE configurations that are too wide to contract end to end
("report per-slice numbers on a seeded sample of slices and say so").

    python tools/sampled_slices.py --config c2      # 3-SAT n=120 m=511, counting semiring
    python tools/sampled_slices.py --config c4      # 3-SAT n=200 m=852, boolean semiring
    python tools/sampled_slices.py --config c5      # 16x16 Wallace multiplier, product fixed to 65521*65519
    torchrun --nproc-per-node 8 ... tools/sampled_slices.py --config c4   # slices split over ranks

The tree (elimination order, 1 trial, sliced to 2^30-entry intermediates) comes from the committed plan files
bench_plans/subcube_*.json (tools/make_subcube_plans.py); all but the last `--keep` sliced wires are pinned on the
host ALONG A KNOWN SOLUTION of the instance (a committed model / the factors 65521 x 65519; `--random-pins` pins
seeded random values instead), and the remaining 2^keep slices are executed exactly like any other plan.  The
printed result is the exact answer for that sub-cube of the assignment space ONLY and is checked against the
independent exact counter (oracle/sharp_sat.py) where that finishes in seconds.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4", choices=["c2", "c4", "c5"])
    ap.add_argument("--target", type=int, default=30)
    ap.add_argument("--keep", type=int, default=4)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--profile", action="store_true", help="time every launch of one slice and print the slowest ops")
    ap.add_argument("--random-pins", action="store_true", help="pin seeded random values (round-1 behaviour: an empty sub-cube)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from tensorsat_b200 import _native
    from tensorsat_b200.distributed import all_reduce_result, slice_block
    from tensorsat_b200.planner import subcube

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    t0 = time.time()
    f = subcube.load(args.config)
    full, net, free, name, mode_name = f.tree, f.tree.network, f.wire_is_free, f.description, f.mode
    sub, subleaves, assignment = f.subcube(args.keep, witness=not args.random_pins, seed=args.seed)
    plan_time = time.time() - t0
    mode = _native.MODE_COUNT if mode_name == "count" else _native.MODE_SAT
    plan = _native.Plan(sub.network.inputs, sub.network.output, sub.network.size_dict, sub.ssa_path, sub.sliced_wires,
                        mode=mode, wire_is_free=free)
    dev_leaves = [_native.upload(t.astype(np.uint8)) for t in subleaves]
    out = _native.empty(plan.info["out_numel"], mode)
    n = plan.info["n_slices"]
    lo, hi = slice_block(n, rank, world)

    if args.profile:
        desc0 = plan.describe()
        kinds = {}
        for o in desc0["ops"]:
            k = (o["kind"], o["per_slice"], o["post"], o["dtype"])
            kinds[k] = kinds.get(k, 0) + 1
        print("op kinds (kind, per_slice, post, dtype): count", kinds, flush=True)
        big = sorted(desc0["ops"], key=lambda o: -o["ops"])[:12]
        for o in big:
            print("largest by ops:", {k: o[k] for k in ("kind", "per_slice", "post", "dtype", "out_bits", "b", "m", "n", "k", "bytes", "ops")}, flush=True)
        plan.execute(dev_leaves, out, 0, n)   # warm-up: loads every kernel
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        plan.execute(dev_leaves, out, 0, n)
        torch.cuda.synchronize()
        print(f"untimed full execution ({n} slices): {(time.perf_counter() - t1) * 1e3:.2f} ms wall")
        t1 = time.perf_counter()
        plan.execute(dev_leaves, out, 0, 1)
        torch.cuda.synchronize()
        print(f"untimed execution of 1 slice: {(time.perf_counter() - t1) * 1e3:.2f} ms wall")
        plan.set_timing(True)
        plan.execute(dev_leaves, out, 0, 1)
        torch.cuda.synchronize()
        desc = plan.describe()
        agg = {}
        for o in desc["ops"]:
            k = (o["kind"], "slice" if o["per_slice"] else ("post" if o["post"] else "pre"))
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += o["ms"]
        agg[("small_batch", "pre+post")] = [sum(1 for l in desc["launches"] if l["op"] < 0), sum(l["ms"] for l in desc["launches"] if l["op"] < 0)]
        for k, (cnt, ms_) in sorted(agg.items()):
            print(f"  {k}: launches={cnt} ms={ms_:.3f}")
        ops = sorted(desc["ops"], key=lambda o: -o["ms"])[:25]
        tot = sum(o["ms"] for o in desc["ops"]) + sum(l["ms"] for l in desc["launches"] if l["op"] < 0)
        print(f"one slice, timed launches: {tot:.3f} ms")
        for o in ops:
            gb = o["bytes"] / (o["ms"] * 1e-3) / 1e9 if o["ms"] else 0
            print(f"{o['kind']:8s} ps={o['per_slice']} post={o['post']} dt={o['dtype']} out_bits={o['out_bits']:2d} b={o['b']:2d} m={o['m']:2d} "
                  f"n={o['n']:2d} k={o['k']:2d} nib=({o['nibL']},{o['nibR']}) ms={o['ms']:9.3f} GB/s={gb:8.1f} intensity={o['ops']/max(o['bytes'],1):7.1f}")
        return

    def step():
        plan.execute(dev_leaves, out, lo, hi)
        if world > 1:
            all_reduce_result(out, mode)

    step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    if rank == 0:
        peak = 6650.0
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peak = float(json.load(fh)["hbm_gbs"])
        except Exception:
            pass
        alg = plan.info["alg_bytes_shared"] + plan.info["alg_bytes_per_slice"] * n
        expected = None
        if args.config in ("c2", "c4"):
            from oracle.sharp_sat import count_models
            from tensorsat_b200.lib.sat import CNFInstance

            cnf = CNFInstance.random(3, f.n_vars, 511 if args.config == "c2" else 852, rng=0)
            expected = count_models(f.n_vars, cnf.clauses, assignment)
            expected = expected if mode_name == "count" else int(expected > 0)
        elif not args.random_pins:
            expected = 1
        print(json.dumps({
            "config": args.config, "instance": name, "mode": mode_name, "n_gpus": world,
            "leaves": net.num_leaves, "wires": net.num_wires, "tree_log2_width": f.extra["tree_log2_width"],
            "sliced_wires_total": len(full.sliced_wires), "log2_slices_full_instance": len(full.sliced_wires),
            "slices_executed": int(n), "sampled": f"{len(assignment)} sliced wires pinned " + (f"to seeded random values (seed {args.seed})" if args.random_pins else "along a committed solution of the instance"),
            "ms_total": ms, "ms_per_slice": ms / n * 1.0, "alg_bytes_executed": alg, "GBps": alg / (ms * 1e-3) / 1e9,
            "frac_hbm_per_gpu": alg / (ms * 1e-3) / 1e9 / peak / world, "peak_arena_bytes": int(plan.info["peak_bytes"]),
            "kernels_once": plan.info["n_kernels_shared"], "kernels_per_slice": plan.info["n_kernels_per_slice"],
            "subcube_result": int(out.cpu().numpy().reshape(-1)[0]), "subcube_expected_by_exact_counter": expected, "host_load_s": round(plan_time, 1),
            "extrapolated_full_instance_s": (ms / n) * 1e-3 * 2.0 ** len(full.sliced_wires),
        }))
        assert expected is None or int(out.cpu().numpy().reshape(-1)[0]) == expected
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
