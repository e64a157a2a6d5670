#!/bin/bash
# round 2: graph capture on a plan-owned stream, footprint-chosen tile wires (A/B), wide-config tests
O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > $O/r2d_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2d_pytest.log; tail -40 $O/r2d_pytest.log
for v in "base:" "tile:TSB_FUSE_TILE=1" "nograph:TSB_GRAPH=0" "tile_nograph:TSB_FUSE_TILE=1 TSB_GRAPH=0"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 400 python bench.py --steps 50 --no-cpu-baseline > $O/r2d_bench_$name.json 2> $O/r2d_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2d_bench_$name.json").read().strip().splitlines()[-1])
    c4 = d["config"].get("c4_subcube") or {}
    print("$name", "ms", round(d["ms_per_step"], 4), "b2b", round(d["config"]["ms_per_step_back_to_back_no_flush"], 4), "e2e_ms", round(d["e2e"]["value"] * 1e3, 4), "graph", d["config"]["cuda_graph"], d["config"]["verdict"], "roof", round(d["roofline"]["frac"], 3), "c4_ms", round(c4.get("value", 0) * 1e3, 2), c4.get("result"))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2d_bench_$name.err").read()[-1500:])
PY
done
for c in c2 c4 c5; do timeout 300 python tools/sampled_slices.py --config $c --keep 4 > $O/r2d_sampled_$c.json 2> $O/r2d_sampled_$c.err; tail -c 700 $O/r2d_sampled_$c.json; tail -3 $O/r2d_sampled_$c.err; done
