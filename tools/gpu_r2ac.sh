#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "contract2_bits or solution_tensors or wide or c4" > $O/r2ac_pytest.log 2>&1; tail -2 $O/r2ac_pytest.log
timeout 600 python tools/sampled_slices.py --config c4 --keep 6 --profile > $O/r2ac_c4_profile.txt 2>&1; grep -A8 "one slice, timed" $O/r2ac_c4_profile.txt; grep "ms_per_slice\|subcube_result" $O/r2ac_c4_profile.txt | head
timeout 600 python bench.py --steps 20 --no-cpu-baseline --no-micro > $O/r2ac_bench.json 2> $O/r2ac_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ac_bench.json").read().strip().splitlines()[-1])
c4 = d["config"]["c4_subcube"]
print("c1 ms %.4f e2e %.4f | c4 %.2f ms, %.4f ms/slice, kernels/slice %d, result %s" % (d["ms_per_step"], d["e2e"]["value"] * 1e3, c4["value"] * 1e3, c4["ms_per_slice"], c4["kernels_per_slice"], c4["result"]))
PY
