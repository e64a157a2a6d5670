#!/usr/bin/env python
"""Summarise ncu artefacts into small text/CSV files that can be committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/r1_launches.csv  > profiles/r1_bench_launches_summary.txt
    python tools/ncu_summary.py report   gpurun_out/x.ncu-rep        > profiles/x_summary.csv
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
]


def launches(path):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"].split("(")[0], float(r["Metric Value"]) * (1e-3 if r["Metric Unit"] == "ns" else 1.0), r["Grid Size"]))
    tot = sum(t for _, t, _ in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for k, t, _ in rows:
        agg[k][0] += 1
        agg[k][1] += t
    print(f"# {path}: {len(rows)} launches, {tot/1e3:.3f} ms summed device time (ncu: cold caches, serialised launches)")
    print("kernel,launches,total_us,share")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k},{n},{t:.1f},{t/tot:.4f}")
    print("\n# 25 longest launches")
    for k, t, g in sorted(rows, key=lambda r: -r[1])[:25]:
        print(f"{k},{t:.1f} us,grid={g}")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = ["Kernel Name"] + [k for k in KEYS if k in idx]
    print(",".join(c + (f" [{units[idx[c]]}]" if units[idx[c]] else "") for c in cols))
    for r in rows[2:]:
        print(",".join('"' + r[idx[c]].split("(")[0] + '"' if c == "Kernel Name" else r[idx[c]] for c in cols))


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
