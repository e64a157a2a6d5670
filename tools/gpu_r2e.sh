#!/bin/bash
# round 2: phase-timed roofline, micro rooflines, graph branch A/B, c5 sampled slices, reference arm
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "graph or streams or arena" > $O/r2e_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2e_pytest.log; tail -5 $O/r2e_pytest.log
for v in "s4:" "s0:TSB_GRAPH_STREAMS=0" "s8:TSB_GRAPH_STREAMS=8"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 400 python bench.py --steps 50 --no-cpu-baseline > $O/r2e_bench_$name.json 2> $O/r2e_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2e_bench_$name.json").read().strip().splitlines()[-1])
    c4 = d["config"].get("c4_subcube") or {}
    r = d["roofline"]
    print("$name", "ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["value"] * 1e3, 4), "graph", d["config"]["cuda_graph"], "roof", round(r["frac"], 3), "phase", {k: round(v, 4) for k, v in r["phase_ms"].items()}, "l1frac", round(r["actual_bound"]["frac"], 3), "c4_ms", round(c4.get("value", 0) * 1e3, 2))
    if "$name" == "s4":
        print("permute", round(d["roofline_permute"]["frac"], 3), [(c["case"][:14], round(c["GBps"])) for c in d["roofline_permute"]["cases"]])
        print("gemm", round(d["roofline_gemm"]["frac"], 3), [(c["b"], c["m"], c["n"], c["k"], c["epilogue"][:5], round(c["TOPs"])) for c in d["roofline_gemm"]["cases"]])
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2e_bench_$name.err").read()[-1500:])
PY
done
timeout 300 python tools/sampled_slices.py --config c5 --keep 4 > $O/r2e_sampled_c5.json 2> $O/r2e_sampled_c5.err; tail -c 700 $O/r2e_sampled_c5.json; tail -3 $O/r2e_sampled_c5.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/r2e_reference.json 2> $O/r2e_reference.err; tail -c 600 $O/r2e_reference.json; tail -4 $O/r2e_reference.err
