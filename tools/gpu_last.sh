#!/bin/bash
O=gpurun_out; mkdir -p $O
for v in 4 0; do
  TSB_WAVE_STREAMS=$v timeout 50 python bench.py --steps 20 --no-cpu-baseline > $O/r1k_streams$v.json 2> $O/r1k_streams$v.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/r1k_streams$v.json").read().splitlines() if l.startswith("{")][-1])
    print("streams=$v ms", round(d["ms_per_step"],4), "b2b", round(d["config"]["ms_per_step_back_to_back_no_flush"],4), "e2e", round(d["e2e"]["value"]*1e3,4), d["config"]["verdict"])
except Exception as e:
    print("streams=$v failed", e)
PY
done
timeout 60 python -m pytest tests/test_gpu_parity.py -x -q -k "slice_blocks or first_model or kissat or partial_outputs" > $O/r1k_pytest_subset.log 2>&1; tail -1 $O/r1k_pytest_subset.log
