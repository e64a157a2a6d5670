#!/bin/bash
O=gpurun_out; mkdir -p $O
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-micro --c4-keep 8 > $O/r2ai_$name.json 2> $O/r2ai_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2ai_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]; c4 = d["config"]["c4_subcube"]
    print("%-10s ms %.4f e2e %.4f loop %.4f verdict %s | c4 %.2f ms %s" % ("$name", d["ms_per_step"], d["e2e"]["value"] * 1e3, r["phase_ms"]["slice_loop"], d["config"]["verdict"], c4["value"] * 1e3, c4["result"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2ai_$name.err").read()[-1500:])
PY
}
run base TSB_X=0
run h7 TSB_LIB=/root/repo/tensorsat_b200/csrc/libtsb_h7.so
run base2 TSB_X=1
run h7b TSB_LIB=/root/repo/tensorsat_b200/csrc/libtsb_h7.so
