#!/bin/bash
# round 2: L1 prefetch of the next vector's fetches in the fused kernel (A/B)
O=gpurun_out; mkdir -p $O
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2n_$name.json 2> $O/r2n_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2n_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-14s ms %.4f e2e %.4f phases %s verdict %s l1frac %.3f hbmfrac %.3f" % ("$name", d["ms_per_step"], d["e2e"]["value"] * 1e3, {k: round(v, 4) for k, v in r["phase_ms"].items()}, d["config"]["verdict"], r["actual_bound"]["frac"], r["frac"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2n_$name.err").read()[-800:])
PY
}
run base TSB_X=0
run pf1 TSB_BITSF_PREFETCH=1
run pf3 TSB_BITSF_PREFETCH=3
run pf15 TSB_BITSF_PREFETCH=15
run pf2 TSB_BITSF_PREFETCH=2
run pf1_np0 TSB_BITSF_PREFETCH=1 TSB_FUSE_PAIR=0
run pf15_np0 TSB_BITSF_PREFETCH=15 TSB_FUSE_PAIR=0
