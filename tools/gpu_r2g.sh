#!/bin/bash
# round 2: candidate plans from the fused-cost search (A/B), GPU test of the native front end, N=1 full bench line
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "native_front or graph" > $O/r2g_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2g_pytest.log; tail -4 $O/r2g_pytest.log
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2g_$name.json 2> $O/r2g_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2g_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-12s ms %.4f e2e %.4f phases %s slices %d launches/step %d verdict %s l1frac %.3f" % ("$name", d["ms_per_step"], d["e2e"]["value"] * 1e3, {k: round(v, 4) for k, v in r["phase_ms"].items()}, d["config"]["n_slices"], d["config"]["kernels_per_step"], d["config"]["verdict"], r["actual_bound"]["frac"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2g_$name.err").read()[-800:])
PY
}
run base TSB_X=0
for i in 0 1 2 3; do run cand$i TSB_BENCH_PLAN=bench_plans/cand_r2_$i.json; done
run base_tile9 TSB_BITSF_TILE_BITS=9
run cand0_tile9 TSB_BENCH_PLAN=bench_plans/cand_r2_0.json TSB_BITSF_TILE_BITS=9
timeout 600 python bench.py --steps 50 > $O/r2g_bench_full.json 2> $O/r2g_bench_full.err; tail -c 3000 $O/r2g_bench_full.json; tail -3 $O/r2g_bench_full.err
