#!/bin/bash
# round 2: L2 eviction-priority hints on the per-slice intermediates
O=gpurun_out; mkdir -p $O
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2v_$name.json 2> $O/r2v_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2v_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-14s ms %.4f e2e %.4f phases %s verdict %s hbmfrac %.3f" % ("$name", d["ms_per_step"], d["e2e"]["value"] * 1e3, {k: round(v, 4) for k, v in r["phase_ms"].items()}, d["config"]["verdict"], r["frac"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2v_$name.err").read()[-1500:])
PY
}
run hint0 TSB_X=0
run hint1 TSB_LIB=/root/repo/tensorsat_b200/csrc/libtsb_h1.so
run hint2 TSB_LIB=/root/repo/tensorsat_b200/csrc/libtsb_h2.so
run hint3 TSB_LIB=/root/repo/tensorsat_b200/csrc/libtsb_h3.so
run hint0b TSB_X=1
