#!/bin/bash
# round 2: candidate trees ranked by the plan compiler's cost under the private-streaming kernels
O=gpurun_out; mkdir -p $O
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2aj_$name.json 2> $O/r2aj_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2aj_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-10s ms %.4f e2e %.4f phases %s verdict %s slices %s" % ("$name", d["ms_per_step"], d["e2e"]["value"] * 1e3, {k: round(v, 4) for k, v in r["phase_ms"].items()}, d["config"]["verdict"], d["config"].get("n_slices")))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2aj_$name.err").read()[-1500:])
PY
}
run base TSB_X=0
for i in 0 1 2 3 4; do run cand$i TSB_BENCH_PLAN=bench_plans/cand_r2b_$i.json; done
run base2 TSB_X=1
