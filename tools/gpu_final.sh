#!/bin/bash
# last visit of the round: A/B of the lane-wire choice, then the evidence of the final state
O=gpurun_out; mkdir -p $O
for v in 1 0; do
  TSB_FUSE_LANES=$v timeout 120 python bench.py --steps 40 --no-cpu-baseline > $O/r1h_lanes$v.json 2> $O/r1h_lanes$v.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/r1h_lanes$v.json").read().splitlines() if l.startswith("{")][-1])
    print("lanes=$v ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]*1e3,4), "frac", round(d["roofline"]["frac"],3), "avg launch us", round(d["roofline"]["avg_launch_ms"]*1e3,2))
except Exception as e:
    print("lanes=$v failed", e)
PY
done
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "bit_packed or assignment or slice_blocks" > $O/r1h_pytest_subset.log 2>&1; tail -1 $O/r1h_pytest_subset.log
timeout 200 python tools/profile_plan.py --top 14 > $O/r1h_profile_plan_sat.txt 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r1h_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r1h_bench_under_ncu.log 2>&1
timeout 200 python bench.py > $O/r1h_bench_n1.json 2> $O/r1h_bench_n1.err; tail -c 300 $O/r1h_bench_n1.json
