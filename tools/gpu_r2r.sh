#!/bin/bash
# round 2: private streaming is the default — ncu capture of the per-slice launches, occupancy variant A/B
O=gpurun_out; mkdir -p $O
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2r_$name.json 2> $O/r2r_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2r_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-14s ms %.4f e2e %.4f phases %s verdict %s hbmfrac %.3f" % ("$name", d["ms_per_step"], d["e2e"]["value"] * 1e3, {k: round(v, 4) for k, v in r["phase_ms"].items()}, d["config"]["verdict"], r["frac"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$O/r2r_$name.err").read()[-1500:])
PY
}
run default TSB_X=0
run c3 TSB_LIB=/root/repo/tensorsat_b200/csrc/libtsb_c3.so TSB_BITSF_CTAS=3
TSB_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_merge_bits_f -s 3 -c 8 -o /tmp/r2r_f -f python tools/profile_plan.py --slices 3 > $O/r2r_ncu.log 2>&1
ncu -i /tmp/r2r_f.ncu-rep --page raw --csv > $O/r2r_bits_f_full.csv 2>/dev/null
for k in 0 1 2 3; do ncu -i /tmp/r2r_f.ncu-rep --page details --kernel-id :::$((k+1)) > $O/r2r_bits_f_details_k$k.txt 2>/dev/null; done
ncu -i /tmp/r2r_f.ncu-rep --page source --csv --kernel-id :::1 > $O/r2r_bits_f_source_k0.csv 2>/dev/null
timeout 300 python tools/profile_plan.py --top 12 > $O/r2r_profile_plan.txt 2>&1; tail -30 $O/r2r_profile_plan.txt
