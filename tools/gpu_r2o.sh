#!/bin/bash
# round 2: byte-address arithmetic in the fused kernel (A/B of two library builds)
O=gpurun_out; mkdir -p $O
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2o_$name.json 2> $O/r2o_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2o_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-14s ms %.4f e2e %.4f phases %s verdict %s l1frac %.3f hbmfrac %.3f" % ("$name", d["ms_per_step"], d["e2e"]["value"] * 1e3, {k: round(v, 4) for k, v in r["phase_ms"].items()}, d["config"]["verdict"], r["actual_bound"]["frac"], r["frac"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2o_$name.err").read()[-800:])
PY
}
run newaddr TSB_X=0
run plainaddr TSB_LIB=/root/repo/tensorsat_b200/csrc/libtsb_plainaddr.so
run newaddr2 TSB_X=1
run plainaddr2 TSB_LIB=/root/repo/tensorsat_b200/csrc/libtsb_plainaddr.so
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "partial_outputs or solution_tensors or full_size or assignment" > $O/r2o_pytest.log 2>&1; tail -3 $O/r2o_pytest.log
