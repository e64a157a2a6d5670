#!/bin/bash
O=gpurun_out; mkdir -p $O
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2an_$name.json 2> $O/r2an_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2an_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-10s ms %.4f e2e %.4f phases %s verdict %s launches/step %d" % ("$name", d["ms_per_step"], d["e2e"]["value"] * 1e3, {k: round(v, 4) for k, v in r["phase_ms"].items()}, d["config"]["verdict"], d["gpu_launches"] // d["steps"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2an_$name.err").read()[-1500:])
PY
}
run m4096 TSB_MERGE_MIN_OUT=4096
run m16384 TSB_MERGE_MIN_OUT=16384
run m32768 TSB_MERGE_MIN_OUT=32768
run m4096b TSB_MERGE_MIN_OUT=4096
run m16384b TSB_MERGE_MIN_OUT=16384
