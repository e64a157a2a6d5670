#!/bin/bash
# round 2: A/B batch on the n=50 bench (slice-loop time from the phase events)
O=gpurun_out; mkdir -p $O
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2f_$name.json 2> $O/r2f_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2f_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-22s ms %.4f e2e %.4f phases %s launches/step %d verdict %s" % ("$name", d["ms_per_step"], d["e2e"]["value"] * 1e3, {k: round(v, 4) for k, v in r["phase_ms"].items()}, d["config"]["kernels_per_step"], d["config"]["verdict"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2f_$name.err").read()[-800:])
PY
}
run base TSB_X=0
run ctas1 TSB_BITSF_CTAS=1
run ctas3 TSB_BITSF_CTAS=3
run ctas4 TSB_BITSF_CTAS=4
run carve0 TSB_BITSF_CARVEOUT=0
run carve25 TSB_BITSF_CARVEOUT=25
run carve100 TSB_BITSF_CARVEOUT=100
run tile8 TSB_BITSF_TILE_BITS=8
run tile9 TSB_BITSF_TILE_BITS=9
run pair0 TSB_FUSE_PAIR=0
run pair0_ctas4 TSB_FUSE_PAIR=0 TSB_BITSF_CTAS=4
run fusemin16 TSB_FUSE_MIN_LOG2=16
run fusemin0 TSB_FUSE_MIN_LOG2=0
run fusemin0_k3 TSB_FUSE_MIN_LOG2=0 TSB_FUSE_MAX_K=3
run chain4 TSB_CHAIN_MAX_OPS=4
run gstreams0 TSB_GRAPH_STREAMS=0
run gstreams8 TSB_GRAPH_STREAMS=8
