#!/bin/bash
# round 2: full GPU suite + smoke on the private-streaming build, then a bench line
O=gpurun_out; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -x > $O/r2w_pytest_gpu.log 2>&1; tail -5 $O/r2w_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2w_smoke.log 2>&1; tail -3 $O/r2w_smoke.log
timeout 900 python bench.py --steps 30 --no-cpu-baseline > $O/r2w_bench_n1.json 2> $O/r2w_bench_n1.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2w_bench_n1.json").read().strip().splitlines()[-1])
print("ms", d["ms_per_step"], "e2e", d["e2e"], "roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "phase_ms") if k in d["roofline"]})
print("c4", {k: d["config"]["c4_subcube"][k] for k in ("value", "ms_per_slice", "result")})
print("permute", d.get("roofline_permute", {}).get("frac"), "gemm", d.get("roofline_gemm", {}).get("frac"))
PY
