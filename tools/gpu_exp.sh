#!/bin/bash
# A/B experiments on the bench plan (one GPU-box visit).  Usage: gpurun --timeout 900 -- 'bash tools/gpu_exp.sh'
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/exp_pytest.log 2>&1; tail -2 $O/exp_pytest.log
run() { # name, args, env...
  name=$1; shift; args=$1; shift
  env "$@" timeout 200 python bench.py --steps 30 --no-cpu-baseline $args > $O/exp_$name.json 2> $O/exp_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$O/exp_$name.json"))
    print("$name", "ms", round(d["ms_per_step"],3), "b2b", round(d["config"]["ms_per_step_back_to_back_no_flush"],3), "e2e", round(d["e2e"]["value"]*1e3,3), "kern", d["config"]["kernels_per_step"], "frac", round(d["roofline"]["frac"],3), "verdict", d["config"]["verdict"])
except Exception as e:
    print("$name failed", e)
PY
}
run pdl1 "" TSB_PDL=1
run pdl0 "" TSB_PDL=0
run pdl1_again "" TSB_PDL=1
run pdl1_count "" TSB_PDL=1
timeout 200 python tools/profile_plan.py --count --top 12 > $O/exp_profile_count.txt 2>&1
