#!/bin/bash
# round 2: two slice lanes (slices alternate between two branches of the captured graph)
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_wide_configs.py tests/test_gpu_parity.py -m gpu -q -x -k "wide or c4 or c2 or c5 or full_size or arena" > $O/r2ao_pytest.log 2>&1; tail -2 $O/r2ao_pytest.log
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-micro --c4-keep 8 > $O/r2ao_$name.json 2> $O/r2ao_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2ao_$name.json").read().strip().splitlines()[-1])
    c4 = d["config"]["c4_subcube"]
    print("%-8s ms %.4f e2e %.4f b2b %.4f verdict %s | c4 %.2f ms %s" % ("$name", d["ms_per_step"], d["e2e"]["value"] * 1e3, d["config"]["ms_per_step_back_to_back_no_flush"], d["config"]["verdict"], c4["value"] * 1e3, c4["result"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2ao_$name.err").read()[-1500:])
PY
}
run auto TSB_X=0
run off TSB_SLICE_LANES=1
run auto2 TSB_X=1
