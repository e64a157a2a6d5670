#!/bin/bash
# round 2: precomputed offset tables of the fused kernel (A/B)
O=gpurun_out; mkdir -p $O
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2j_$name.json 2> $O/r2j_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2j_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-14s ms %.4f e2e %.4f phases %s verdict %s l1frac %.3f hbmfrac %.3f" % ("$name", d["ms_per_step"], d["e2e"]["value"] * 1e3, {k: round(v, 4) for k, v in r["phase_ms"].items()}, d["config"]["verdict"], r["actual_bound"]["frac"], r["frac"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2j_$name.err").read()[-800:])
PY
}
run tables TSB_X=0
run notables TSB_BITSF_TABLES=0
run tables_tile9 TSB_BITSF_TILE_BITS=9
run tables_pair2 TSB_FUSE_PAIR=2
run tables_pair2_t9 TSB_FUSE_PAIR=2 TSB_BITSF_TILE_BITS=9
run tables_ctas3 TSB_BITSF_CTAS=3
run tables_pair0 TSB_FUSE_PAIR=0
run tables_cand0 TSB_BENCH_PLAN=bench_plans/cand_r2_0.json
