#!/bin/bash
# round 2: private streaming (two / three pair wires held by slot 0 only) against the one-pair-wire kernel
O=gpurun_out; mkdir -p $O
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2q_$name.json 2> $O/r2q_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2q_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-14s ms %.4f e2e %.4f phases %s verdict %s hbmfrac %.3f" % ("$name", d["ms_per_step"], d["e2e"]["value"] * 1e3, {k: round(v, 4) for k, v in r["phase_ms"].items()}, d["config"]["verdict"], r["frac"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2q_$name.err").read()[-1500:])
PY
}
run pair1 TSB_FUSE_PAIR=1
run pair3 TSB_FUSE_PAIR=3
run pair2 TSB_FUSE_PAIR=2
run pair3c3 TSB_FUSE_PAIR=3 TSB_BITSF_CTAS=3
run pair3t9 TSB_FUSE_PAIR=3 TSB_BITSF_TILE_BITS=9
run pair1b TSB_FUSE_PAIR=1
TSB_FUSE_PAIR=3 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "partial_outputs or solution_tensors or full_size or assignment" > $O/r2q_pytest.log 2>&1; tail -3 $O/r2q_pytest.log
