#!/bin/bash
# round 2: fusing the small per-slice chain of the config-4 plan (TSB_FUSE_MIN_LOG2)
O=gpurun_out; mkdir -p $O
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --no-cpu-baseline --no-micro --c4-keep 8 > $O/r2x_$name.json 2> $O/r2x_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2x_$name.json").read().strip().splitlines()[-1])
    c4 = d["config"]["c4_subcube"]
    print("%-10s c1 ms %.4f | c4 %.2f ms, %.4f ms/slice, kernels/slice %d, once %d, result %s" % ("$name", d["ms_per_step"], c4["value"] * 1e3, c4["ms_per_slice"], c4["kernels_per_slice"], c4["kernels_once"], c4["result"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2x_$name.err").read()[-1500:])
PY
}
run f22 TSB_X=0
run f18 TSB_FUSE_MIN_LOG2=18
run f14 TSB_FUSE_MIN_LOG2=14
run f0 TSB_FUSE_MIN_LOG2=0
