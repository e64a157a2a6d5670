#!/bin/bash
# round 2: last check of the committed state — full GPU suite, smoke, one bench line
O=gpurun_out; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -x > $O/r2ak_pytest_gpu.log 2>&1; tail -3 $O/r2ak_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2ak_smoke.log 2>&1; tail -1 $O/r2ak_smoke.log
timeout 900 python bench.py --steps 50 --no-cpu-baseline > $O/r2ak_bench_n1.json 2> $O/r2ak_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ak_bench_n1.json").read().strip().splitlines()[-1])
print("N=1 ms", d["ms_per_step"], "e2e", d["e2e"]["value"] * 1e3, "roofline frac", d["roofline"]["frac"], "c4", d["config"]["c4_subcube"]["value"], d["config"]["c4_subcube"]["result"])
PY
