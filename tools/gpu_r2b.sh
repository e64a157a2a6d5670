#!/bin/bash
# round 2: CUDA-graph replay — GPU suite, then A/B bench lines
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2b_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2b_pytest.log; tail -15 $O/r2b_pytest.log
for v in "graph:" "nograph:TSB_GRAPH=0" "graph_s0:TSB_GRAPH_STREAMS=0" "graph_s8:TSB_GRAPH_STREAMS=8"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python bench.py --steps 50 --no-cpu-baseline > $O/r2b_bench_$name.json 2> $O/r2b_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2b_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"], 4), "b2b", round(d["config"]["ms_per_step_back_to_back_no_flush"], 4), "e2e_ms", round(d["e2e"]["value"] * 1e3, 4), "launches", d["gpu_launches"], d["config"]["verdict"])
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2b_bench_$name.err").read()[-1500:])
PY
done
