#!/bin/bash
# round 2: A/B of the four-vector (two pair wires) fused kernel and of tile mode 2
O=gpurun_out; mkdir -p $O
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2i_$name.json 2> $O/r2i_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2i_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-14s ms %.4f e2e %.4f phases %s verdict %s l1frac %.3f hbmfrac %.3f" % ("$name", d["ms_per_step"], d["e2e"]["value"] * 1e3, {k: round(v, 4) for k, v in r["phase_ms"].items()}, d["config"]["verdict"], r["actual_bound"]["frac"], r["frac"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2i_$name.err").read()[-800:])
PY
}
run base TSB_X=0
run pair2 TSB_FUSE_PAIR=2
run pair2_ctas2 TSB_FUSE_PAIR=2 TSB_BITSF_CTAS=2
run pair2_ctas4 TSB_FUSE_PAIR=2 TSB_BITSF_CTAS=4
run pair2_tile9 TSB_FUSE_PAIR=2 TSB_BITSF_TILE_BITS=9
run tile2 TSB_FUSE_TILE=2
run pair2_tile2 TSB_FUSE_PAIR=2 TSB_FUSE_TILE=2
run pair2_cand0 TSB_FUSE_PAIR=2 TSB_BENCH_PLAN=bench_plans/cand_r2_0.json
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "partial_outputs or solution_tensors or assignment_search or full_size" > $O/r2i_pytest.log 2>&1; tail -3 $O/r2i_pytest.log
