#!/bin/bash
# round 2, closing run at one GPU: full GPU suite, smoke, the bench line as the driver runs it, the reference arm
O=gpurun_out; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -x > $O/r2ae_pytest_gpu.log 2>&1; tail -3 $O/r2ae_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2ae_smoke.log 2>&1; tail -2 $O/r2ae_smoke.log
timeout 900 python bench.py > $O/r2ae_bench_n1.json 2> $O/r2ae_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ae_bench_n1.json").read().strip().splitlines()[-1])
print("N=1 steps", d["steps"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"] * 1e3, "cpu", d["cpu_baseline"]["value"], "roofline frac", d["roofline"]["frac"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print("c4", {k: d["config"]["c4_subcube"][k] for k in ("value", "ms_per_slice", "result", "kernels_per_slice")})
print("permute", d["roofline_permute"]["frac"], "gemm", d["roofline_gemm"]["frac"])
PY
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2ae_reference.json 2> $O/r2ae_reference.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ae_reference.json").read().strip().splitlines()[-1])
print("reference arm", d["value"], d["unit"], d["cpu_baseline"]["sample"][:200])
PY
