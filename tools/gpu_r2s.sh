#!/bin/bash
# round 2: two-sided private streaming; A/B against one-sided and NP=1
O=gpurun_out; mkdir -p $O
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2s_$name.json 2> $O/r2s_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2s_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-14s ms %.4f e2e %.4f phases %s verdict %s hbmfrac %.3f" % ("$name", d["ms_per_step"], d["e2e"]["value"] * 1e3, {k: round(v, 4) for k, v in r["phase_ms"].items()}, d["config"]["verdict"], r["frac"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2s_$name.err").read()[-1500:])
PY
}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "private_streaming or partial_outputs" > $O/r2s_pytest0.log 2>&1; tail -3 $O/r2s_pytest0.log
run two TSB_X=0
run one TSB_BITSF_TWO_SIDED=0
run two1 TSB_BITSF_TWO_SIDED=1
run pair1 TSB_FUSE_PAIR=1
run two_b TSB_X=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "solution_tensors or full_size or assignment" > $O/r2s_pytest.log 2>&1; tail -3 $O/r2s_pytest.log
