#!/bin/bash
O=gpurun_out; mkdir -p $O
TSB_BENCH_PLAN=bench_plans/candidate_c1_8slices.json timeout 40 python bench.py --steps 20 --no-cpu-baseline > $O/r1l_candidate.json 2> $O/r1l_candidate.err
timeout 40 python bench.py --steps 20 --no-cpu-baseline > $O/r1l_current.json 2> $O/r1l_current.err
python - <<PY
import json
for n in ("candidate","current"):
    try:
        d=json.loads([l for l in open("$O/r1l_%s.json"%n).read().splitlines() if l.startswith("{")][-1])
        print(n, "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]*1e3,4), d["config"]["verdict"], d["config"]["n_slices"], d["config"]["kernels_per_step"])
    except Exception as e:
        print(n, "failed", e)
PY
