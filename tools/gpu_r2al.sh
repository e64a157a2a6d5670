#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_abi.py -m gpu -q -x -k "from_host or kissat or frontend or assignment" > $O/r2al_pytest.log 2>&1; tail -2 $O/r2al_pytest.log
for i in 1 2; do
timeout 600 python bench.py --steps 50 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2al_bench$i.json 2> $O/r2al_bench$i.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2al_bench$i.json").read().strip().splitlines()[-1])
print("ms %.4f e2e %.4f" % (d["ms_per_step"], d["e2e"]["value"] * 1e3), d["e2e"])
PY
done
