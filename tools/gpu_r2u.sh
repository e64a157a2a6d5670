#!/bin/bash
# round 2: the same tree sliced one / two wires further (intermediates of a slice fit the L2)
O=gpurun_out; mkdir -p $O
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2u_$name.json 2> $O/r2u_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2u_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-14s ms %.4f e2e %.4f phases %s verdict %s hbmfrac %.3f" % ("$name", d["ms_per_step"], d["e2e"]["value"] * 1e3, {k: round(v, 4) for k, v in r["phase_ms"].items()}, d["config"]["verdict"], r["frac"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2u_$name.err").read()[-1500:])
PY
}
export TSB_BITSF_TWO_SIDED=1
run base TSB_X=0
run s29 TSB_BENCH_PLAN=bench_plans/c1_sliced_29.json
run s29two2 TSB_BENCH_PLAN=bench_plans/c1_sliced_29.json TSB_BITSF_TWO_SIDED=2
run s28 TSB_BENCH_PLAN=bench_plans/c1_sliced_28.json
run s29m17 TSB_BENCH_PLAN=bench_plans/c1_sliced_29.json TSB_BITSF_MIN_ITEMS_LOG2=17
run s29m16 TSB_BENCH_PLAN=bench_plans/c1_sliced_29.json TSB_BITSF_MIN_ITEMS_LOG2=16
run s28m16 TSB_BENCH_PLAN=bench_plans/c1_sliced_28.json TSB_BITSF_MIN_ITEMS_LOG2=16
run base_m17 TSB_BITSF_MIN_ITEMS_LOG2=17
