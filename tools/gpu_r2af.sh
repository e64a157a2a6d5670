#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "from_host or kissat or frontend" > $O/r2af_pytest.log 2>&1; tail -2 $O/r2af_pytest.log
timeout 600 python bench.py --steps 50 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2af_bench.json 2> $O/r2af_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2af_bench.json").read().strip().splitlines()[-1])
print("ms %.4f e2e %.4f" % (d["ms_per_step"], d["e2e"]["value"] * 1e3))
PY
python -c "import tensorsat_b200.contractions.tree as t; print('helper', t._tsb_host)"
