#!/usr/bin/env python
"""
bench.py — contraction-phase benchmark of the B200 executor (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (every N): BASELINE config 1, random 3-SAT n=50 m=213 (ratio 4.26) seed 0, SAT decision
with all inputs unknown, the reference's bintree network (1008 leaf tensors, 795 wires); an
elimination-order contraction tree sliced to 2^30-entry intermediates = 16 slices.  A "step" is
one whole contraction (shared part + all 16 slices + combine); with N GPUs the 16 slices are split
into contiguous blocks, one per rank, and combined by one NCCL all-reduce (max).  Total work is
fixed, so scaling is "strong"; the metric is the contraction time in seconds.

Timed region: leaves resident in HBM, K executions bracketed by barrier + synchronize, one pair of
CUDA events per execution on the launching stream, max over ranks.  The inputs of a contraction are
5 KB of leaf tensors and every intermediate is produced inside the step; the L2 is flushed before
every timed execution anyway (the back-to-back figure is reported next to it).  ``e2e`` times the same contraction through
``TreeContraction.contract_from_host`` with the leaves in host memory (one H2D of the packed
leaves and one D2H of the verdict inside the timed region).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_VARS, N_CLAUSES, SEED, K_SAT = 50, 213, 0, 3
TARGET_LOG2 = 30
PLAN_FILE = os.environ.get("TSB_BENCH_PLAN") or os.path.join(ROOT, "bench_plans", f"c1_n50_m213_seed0_w{TARGET_LOG2}.json")
WORKLOAD = (f"random {K_SAT}-SAT n={N_VARS} m={N_CLAUSES} seed={SEED}, SAT decision (all inputs unknown), "
            f"bintree network, elimination-order tree sliced to 2^{TARGET_LOG2}-entry intermediates")


def load_tree():
    from tensorsat_b200.lib.sat import CNFInstance
    from tensorsat_b200.planner import ContractionTree, Network
    from tensorsat_b200.planner.search import find_tree_elimination

    cnf = CNFInstance.random(K_SAT, N_VARS, N_CLAUSES, rng=SEED)
    diagram = (cnf.inputs(None) >> cnf.diagram()).flatten()
    net = Network.from_diagram(diagram)
    tree = None
    if os.path.exists(PLAN_FILE):
        with open(PLAN_FILE) as f:
            tree = ContractionTree.from_json(json.load(f))
        if tree.network != net:
            tree = None
    if tree is None:
        tree = find_tree_elimination(net, trials=12, seed=0).slice_to(TARGET_LOG2)
        try:
            os.makedirs(os.path.dirname(PLAN_FILE), exist_ok=True)
            with open(PLAN_FILE, "w") as f:
                f.write(tree.dumps())
        except OSError:
            pass
    return cnf, diagram, tree


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.gpu, self.lines, self.proc = gpu_index, [], None
        self.window: tuple[float, float] | None = None   # host-clock bounds of the timed region

    def wait_first_sample(self, timeout: float = 3.0) -> None:
        t0 = time.time()
        while not self.lines and time.time() - t0 < timeout:
            time.sleep(0.01)

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self) -> None:
        assert self.proc and self.proc.stdout
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines, window = self.lines, "run (warm-up + timed region + e2e)"
        if self.window:
            # a sample is stamped when it is read: it describes the ~20 ms before that
            inside = [x for x in self.lines if self.window[0] <= x[0] <= self.window[1] + 0.03]
            if inside:
                lines, window = inside, "timed region"
        for _, ln in lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def hbm_peak() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


_FINE: dict = {}


def cpu_sample(tree, diagram, n_sample: int, target_log2: int = 22) -> dict:
    """The reference arithmetic (numpy oracle) on a bounded sample of the same plan: the tree is sliced
    further so that one slice fits the host comfortably, `n_sample` slices are timed and the whole-job
    time is extrapolated (slices have identical shapes and cost)."""
    from oracle import fin_rel_oracle as O

    key = (id(tree), target_log2)
    if key not in _FINE:   # slicing the tree to host size takes seconds: once per run, not once per step
        _FINE[key] = tree.with_sliced(()).slice_to(target_log2)
    fine = _FINE[key]
    leaves = [b.tensor for b in diagram.boxes]
    n = fine.num_slices
    ids = [int(i * n / n_sample) for i in range(n_sample)]
    net = fine.network
    O.execute_tree(net.inputs, net.output, leaves, fine.ssa_path, fine.sliced_wires, slice_ids=ids[:1], size_dict=net.size_dict)
    t0 = time.perf_counter()
    O.execute_tree(net.inputs, net.output, leaves, fine.ssa_path, fine.sliced_wires, slice_ids=ids, size_dict=net.size_dict)
    dt = time.perf_counter() - t0
    est = dt / n_sample * n
    return {"value": est, "unit": "s", "cores": 1, "kind": "port",
            "host_cores_available": os.cpu_count(),
            "sample": (f"numpy oracle (reference arithmetic, single-threaded integer einsum) on {n_sample} of {n} slices "
                       f"of the same tree sliced to 2^{target_log2}-entry intermediates, {dt:.2f} s measured, "
                       f"extrapolated x{n / n_sample:.0f}")}


def c4_subcube_leg(world: int, rank: int, keep: int, steps: int) -> dict | None:
    """BASELINE config 4 (random 3-SAT n=200 m=852 seed 0, the north-star's sliced multi-GPU instance) on a SAMPLED
    sub-cube: the tree (width 136) is sliced to 2^30-entry intermediates = 2^106 slices; all but `keep` sliced
    wires are pinned along a committed model of the instance, the remaining 2^keep slices are split over the
    ranks in contiguous blocks and combined by one all-reduce (max).  The sub-cube contains the model, so its
    exact answer is SAT (checked)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from tensorsat_b200 import _native
    from tensorsat_b200.distributed import all_reduce_result, slice_block
    from tensorsat_b200.planner import subcube

    if not os.path.exists(subcube.path_of("c4")):
        return None
    f = subcube.load("c4")
    sub, subleaves, asg = f.subcube(keep)
    net = sub.network
    plan = _native.Plan(net.inputs, net.output, net.size_dict, sub.ssa_path, sub.sliced_wires, mode=_native.MODE_SAT)
    leaves = [_native.upload(np.ascontiguousarray(t, dtype=np.uint8)) for t in subleaves]
    out = _native.empty(plan.info["out_numel"], _native.MODE_SAT)
    n = int(plan.info["n_slices"])
    lo, hi = slice_block(n, rank, world)

    def step() -> None:
        plan.execute(leaves, out, lo, hi)
        if world > 1:
            all_reduce_result(out, _native.MODE_SAT)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / steps
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    result = int(out.cpu().numpy().reshape(-1)[0])
    info = plan.info
    alg = info["alg_bytes_shared"] + info["alg_bytes_per_slice"] * n
    res = {
        "workload": (f"{f.description}; elimination-order tree (width {f.extra['tree_log2_width']:.0f}) sliced to 2^30-entry "
                     f"intermediates = 2^{f.extra['n_sliced']} slices of the full instance; SAMPLED: {len(asg)} sliced wires pinned "
                     f"along a committed model, the other 2^{keep} = {n} slices executed"),
        "value": ms * 1e-3, "unit": "s", "ms_per_slice": ms / max(hi - lo, 1), "n_slices": n, "slices_per_rank": int(hi - lo),
        "steps": steps, "scaling": "strong", "result": "SAT" if result else "UNSAT", "expected": "SAT (the sub-cube contains the model)",
        "kernels_once": int(info["n_kernels_shared"]), "kernels_per_slice": int(info["n_kernels_per_slice"]),
        "alg_bytes_moved_per_contraction": alg, "alg_GBps": alg / (ms * 1e-3) / 1e9,
        "peak_arena_bytes": int(info["peak_bytes"]), "timing": "CUDA events per contraction incl. the all-reduce, max over ranks",
    }
    assert result == 1, "the sampled sub-cube of config 4 contains a model: the contraction must say SAT"
    plan.close()
    return res


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cnf, diagram, tree = load_tree()
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_sample(tree, diagram, n_sample=4)
        if i >= args.warmup:
            vals.append(r)
    v = sum(x["value"] for x in vals) / len(vals)
    line = {
        "impl": "reference", "metric": "random 3-SAT contraction time", "value": v, "unit": "s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": v * 1e3,
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_slices": tree.num_slices},
        "cpu_baseline": {**vals[-1], "value": v},
        "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pack", action="store_true", help="keep boolean intermediates at one byte per entry (A/B runs)")
    ap.add_argument("--batch-log2", type=int, default=None, help="override the plan's once-only size limit (experiments)")
    ap.add_argument("--c4-keep", type=int, default=10, help="config-4 sub-cube leg: 2^keep slices (0 < keep; -1: skip the leg)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from tensorsat_b200 import _native
    from tensorsat_b200.contractions.tree import TreeContraction
    from tensorsat_b200.distributed import all_reduce_result, slice_block
    from tensorsat_b200.lang.fin_rel import CudaFinRel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cnf, diagram, tree = load_tree()
    contraction = TreeContraction(CudaFinRel, diagram.wiring, tree=tree, bit_packing=not args.no_pack, batch_log2=-1 if args.batch_log2 is None else args.batch_log2)
    plan = contraction.plan()
    n_slices = plan.info["n_slices"]
    lo, hi = slice_block(n_slices, rank, world)
    leaves_dev = [contraction._leaf_tensor(b) for b in diagram.boxes]
    out = _native.empty(plan.info["out_numel"], contraction.mode)

    def step() -> None:
        plan.execute(leaves_dev, out, lo, hi)
        if world > 1:
            all_reduce_result(out, contraction.mode)

    def sync_all() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
    for _ in range(args.warmup):
        step()
    sync_all()
    # back-to-back (no flush): K executions between one pair of events
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync_all()
    ms_b2b = e0.elapsed_time(e1)
    # the reported number: L2 flushed before every timed step (a 256 MiB fill, twice the 126 MB L2, outside the
    # events), one pair of CUDA events per step on the launching stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = _native.launch_count()
    sync_all()
    t_begin = time.time()
    for a, b in evs:
        flush.fill_(1)
        a.record()
        step()
        b.record()
    sync_all()
    sampler.window = (t_begin, time.time())
    ms = sum(a.elapsed_time(b) for a, b in evs)
    launches = _native.launch_count() - launches0
    verdict = bool(out.cpu().numpy().reshape(-1)[0])
    t = torch.tensor([ms, float(launches), ms_b2b], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, launches, ms_b2b = float(tmax[0]), int(tsum[1]), float(tmax[2])
    ms_per_step = ms / args.steps

    # ---- end to end through the public API with HOST leaf buffers
    leaves_host = [b.tensor for b in diagram.boxes]
    for _ in range(2):
        contraction.contract_from_host(leaves_host, (lo, hi))
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = contraction.contract_from_host(leaves_host, (lo, hi))
        if world > 1:
            r = torch.from_numpy(res.reshape(-1).copy()).cuda()
            all_reduce_result(r, contraction.mode)
            res = r.cpu().numpy()
    sync_all()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt[0])
    clocks = sampler.stop() if rank == 0 else None
    desc_graph = plan.describe()
    # ---- the north-star's sliced instance (config 4) on a sampled sub-cube, same ranks
    c4 = c4_subcube_leg(world, rank, args.c4_keep, steps=3) if args.c4_keep > 0 else None

    # ---- per-kernel roofline: instrumented pass (CUDA events around every launch of the plan)
    roof = None
    if rank == 0:
        plan.set_timing(True)
        plan.execute(leaves_dev, out, lo, hi)
        torch.cuda.synchronize()
        desc = plan.describe()
        plan.set_timing(False)
        peak, peak_src = hbm_peak()
        share = (hi - lo) if n_slices > 1 else 1
        # dominant kernel: the per-slice merge launches (the large middle of the tree); k_merge_bits when the plan
        # runs them bit-packed (bytes = 1/8 per entry, the representation the kernel actually moves), else k_merge_sat
        packed = plan.info["n_packed_steps"] > 0
        kind = "bits" if packed else "merge"
        merge = [o for o in desc["ops"] if o["kind"] == kind and (o["per_slice"] or n_slices == 1)]
        mbytes = sum(o["bytes"] * (share if o["per_slice"] else 1) for o in merge)
        mms = sum(o["ms"] for o in merge)
        all_ms = sum(o["ms"] for o in desc["ops"])
        n_launch = sum((share if o["per_slice"] else 1) for o in merge)
        achieved = mbytes / (mms * 1e-3) / 1e9 if mms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "bits_traffic.json" if packed else "merge_traffic.json")
        if os.path.exists(tp):
            try:
                with open(tp) as f:
                    traffic = json.load(f).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "kernel": ("k_merge_bits_f<NOPS,KB,MASKS,NP> (fused bit-packed steps)" if packed else "k_merge_sat<NIBL,NIBR>") + ", per-slice launches",
                "bytes_counted": ("operands + result of each launch at 1/8 byte per entry; intermediates of the fused steps are not stored and not counted"
                                  if packed else "operands + result of each step at 1 byte per entry"),
                "bytes_per_entry": 0.125 if packed else 1, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": mbytes / max(n_launch, 1), "avg_launch_ms": mms / max(n_launch, 1),
                "launches_per_step": n_launch, "share_of_step": mms / all_ms if all_ms else None,
                "whole_step_alg_GBps": (plan.info["alg_bytes_shared"] + plan.info["alg_bytes_per_slice"] * share)
                / (ms_per_step * 1e-3) / 1e9}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_sample(tree, diagram, n_sample=128)
        line = {
            "metric": "random 3-SAT contraction time", "value": ms_per_step * 1e-3, "unit": "s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_slices": int(n_slices), "slices_per_rank": int(hi - lo),
                       "verdict": "SAT" if verdict else "UNSAT",
                       "l2": "flushed before every timed step (256 MiB fill outside the per-step CUDA events)",
                       "ms_per_step_back_to_back_no_flush": ms_b2b / args.steps,
                       "peak_arena_bytes": int(plan.info["peak_bytes"]), "kernels_per_step": launches // max(args.steps, 1),
                       "alg_bytes_per_step": plan.info["alg_bytes_shared"] + plan.info["alg_bytes_per_slice"] * n_slices,
                       "alg_bytes_per_step_at_1B_per_entry": plan.info["alg_bytes_u8_shared"] + plan.info["alg_bytes_u8_per_slice"] * n_slices,
                       "bit_packed_steps": int(plan.info["n_packed_steps"]), "conversions": int(plan.info["n_conversions"]),
                       "cuda_graph": {k: desc_graph[k] for k in ("graph_captures", "graph_replays")},
                       "c4_subcube": c4},
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e_s, "unit": "s", "h2d_bytes_per_step": contraction.h2d_bytes,
                    "d2h_bytes_per_step": contraction.d2h_bytes},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
