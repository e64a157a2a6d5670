#!/bin/bash
# round 2: the same tree with fewer, larger slices (8 / 4 instead of 16)
O=gpurun_out; mkdir -p $O
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2ah_$name.json 2> $O/r2ah_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2ah_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-10s ms %.4f e2e %.4f phases %s verdict %s hbmfrac %.3f" % ("$name", d["ms_per_step"], d["e2e"]["value"] * 1e3, {k: round(v, 4) for k, v in r["phase_ms"].items()}, d["config"]["verdict"], r["frac"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2ah_$name.err").read()[-1500:])
PY
}
run base TSB_X=0
run s31 TSB_BENCH_PLAN=bench_plans/c1_sliced_31.json
run s32 TSB_BENCH_PLAN=bench_plans/c1_sliced_32.json
run base2 TSB_X=1
