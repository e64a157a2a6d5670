#!/bin/bash
# round 2: staged fused kernel (k_merge_bits_s) — parity first, then A/B against k_merge_bits_f
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "staged" > $O/r2p_pytest_staged.log 2>&1; tail -3 $O/r2p_pytest_staged.log
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2p_$name.json 2> $O/r2p_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2p_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-14s ms %.4f e2e %.4f phases %s verdict %s hbmfrac %.3f" % ("$name", d["ms_per_step"], d["e2e"]["value"] * 1e3, {k: round(v, 4) for k, v in r["phase_ms"].items()}, d["config"]["verdict"], r["frac"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2p_$name.err").read()[-1500:])
PY
}
run staged0 TSB_BITSF_STAGED=0
run staged1 TSB_BITSF_STAGED=1
run t9s4 TSB_BITSF_STAGE_TILE=9 TSB_BITSF_STAGES=4
run t9s2 TSB_BITSF_STAGE_TILE=9 TSB_BITSF_STAGES=2
run t8s4 TSB_BITSF_STAGE_TILE=8 TSB_BITSF_STAGES=4
run staged0b TSB_BITSF_STAGED=0
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "partial_outputs or solution_tensors or full_size or assignment" > $O/r2p_pytest.log 2>&1; tail -3 $O/r2p_pytest.log
