#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python bench.py > $O/r2ap_bench_n1.json 2> $O/r2ap_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ap_bench_n1.json").read().strip().splitlines()[-1])
print("N=1 steps", d["steps"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"] * 1e3, "cpu", d["cpu_baseline"]["value"], "roofline frac", d["roofline"]["frac"], "launches", d["gpu_launches"])
print("c4", {k: d["config"]["c4_subcube"][k] for k in ("value", "ms_per_slice", "result", "kernels_per_slice")})
print("permute", d["roofline_permute"]["frac"], "gemm", d["roofline_gemm"]["frac"])
PY
timeout 300 python tools/sampled_slices.py --config c4 --keep 4 > $O/r2ap_sampled_c4.json 2> $O/r2ap_sampled_c4.err; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2ap_sampled_c4.json").read().strip().splitlines() if l.startswith("{")][-1])
print("c4 keep4", "ms/slice", round(d["ms_per_slice"], 3), "result", d["subcube_result"], "expected", d["subcube_expected_by_exact_counter"])
PY
