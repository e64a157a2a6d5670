#!/bin/bash
# A/B experiments on the bench plan (one GPU-box visit).  Usage: gpurun --timeout 900 -- 'bash tools/gpu_exp.sh'
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "bit_packed or assignment" > $O/exp_pytest.log 2>&1; tail -2 $O/exp_pytest.log
run() { # name, args, env...
  name=$1; shift; args=$1; shift
  env "$@" timeout 200 python bench.py --steps 30 --no-cpu-baseline $args > $O/exp_$name.json 2> $O/exp_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$O/exp_$name.json"))
    print("$name", "ms", round(d["ms_per_step"],3), "b2b", round(d["config"]["ms_per_step_back_to_back_no_flush"],3), "e2e", round(d["e2e"]["value"]*1e3,3), "kern", d["config"]["kernels_per_step"], "frac", round(d["roofline"]["frac"],3), "arena", d["config"]["peak_arena_bytes"]>>20)
except Exception as e:
    print("$name failed", e)
PY
}
run pair0 "" TSB_FUSE_PAIR=0
run pair1 "" TSB_FUSE_PAIR=1
run pair1_b24 "--batch-log2 24" TSB_FUSE_PAIR=1
run pair0_b24 "--batch-log2 24" TSB_FUSE_PAIR=0
run pair0_b26 "--batch-log2 26" TSB_FUSE_PAIR=0
run pair0_fuse20 "" TSB_FUSE_PAIR=0 TSB_FUSE_MIN_LOG2=20
run pair0_fuse24 "" TSB_FUSE_PAIR=0 TSB_FUSE_MIN_LOG2=24
TSB_FUSE_PAIR=1 timeout 200 python tools/profile_plan.py --top 30 > $O/exp_profile_plan_pair1.txt 2>&1
TSB_FUSE_PAIR=0 timeout 200 python tools/profile_plan.py --top 30 > $O/exp_profile_plan_pair0.txt 2>&1
