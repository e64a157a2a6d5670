#!/bin/bash
# round 2, closing numbers at one GPU after the config-4 work: default bench line, wide configs on their sub-cubes
O=gpurun_out; mkdir -p $O
timeout 900 python bench.py > $O/r2ag_bench_n1.json 2> $O/r2ag_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ag_bench_n1.json").read().strip().splitlines()[-1])
print("N=1 steps", d["steps"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"] * 1e3, "cpu", d["cpu_baseline"]["value"], "roofline frac", d["roofline"]["frac"], "launches", d["gpu_launches"])
print("c4", {k: d["config"]["c4_subcube"][k] for k in ("value", "ms_per_slice", "result", "kernels_per_slice")})
print("permute", d["roofline_permute"]["frac"], "gemm", d["roofline_gemm"]["frac"])
PY
for c in c2 c4 c5; do timeout 900 python tools/sampled_slices.py --config $c --keep 4 > $O/r2ag_sampled_$c.json 2> $O/r2ag_sampled_$c.err; python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2ag_sampled_$c.json").read().strip().splitlines() if l.startswith("{")][-1])
print("$c", "ms/slice", round(d["ms_per_slice"], 3), "GB/s", round(d["GBps"], 1), "result", d["subcube_result"], "expected", d["subcube_expected_by_exact_counter"], "kernels/slice", d["kernels_per_slice"])
PY
done
timeout 900 python -m pytest tests/test_gpu_wide_configs.py -m gpu -q -x > $O/r2ag_pytest_wide.log 2>&1; tail -2 $O/r2ag_pytest_wide.log
