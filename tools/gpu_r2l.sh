#!/bin/bash
# round 2: small-op descriptor staging + split batches: GPU suite and bench
O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x > $O/r2l_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2l_pytest.log; tail -4 $O/r2l_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2l_smoke.log 2>&1; tail -2 $O/r2l_smoke.log
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2l_$name.json 2> $O/r2l_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2l_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-14s ms %.4f e2e %.4f phases %s verdict %s l1frac %.3f hbmfrac %.3f" % ("$name", d["ms_per_step"], d["e2e"]["value"] * 1e3, {k: round(v, 4) for k, v in r["phase_ms"].items()}, d["config"]["verdict"], r["actual_bound"]["frac"], r["frac"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2l_$name.err").read()[-800:])
PY
}
run base TSB_X=0
run base2 TSB_X=1
run nograph TSB_GRAPH=0
