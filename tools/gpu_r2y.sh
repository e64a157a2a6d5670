#!/bin/bash
# round 2, final evidence at one GPU: full bench line, reference arm, ncu launch list + full capture of the fused kernel
O=gpurun_out; mkdir -p $O
timeout 900 python bench.py --steps 50 > $O/r2y_bench_n1.json 2> $O/r2y_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2y_bench_n1.json").read().strip().splitlines()[-1])
print("N=1 ms", d["ms_per_step"], "e2e", d["e2e"]["value"] * 1e3, "cpu", d["cpu_baseline"], "roofline frac", d["roofline"]["frac"], "launches", d["gpu_launches"])
print("c4", {k: d["config"]["c4_subcube"][k] for k in ("value", "ms_per_slice", "result", "kernels_per_slice")})
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2y_reference.json 2> $O/r2y_reference.err; tail -c 600 $O/r2y_reference.json
TSB_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r2y_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2y_ncu_bench.log 2>&1
python tools/ncu_summary.py launches $O/r2y_bench_launches.csv > $O/r2y_bench_launches_summary.txt; head -12 $O/r2y_bench_launches_summary.txt
TSB_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_merge_bits_f -s 7 -c 8 -o /tmp/r2y_f -f python tools/profile_plan.py --slices 4 > $O/r2y_ncu.log 2>&1
ncu -i /tmp/r2y_f.ncu-rep --page raw --csv > $O/r2y_bits_f_full.csv 2>/dev/null
for k in 0 1 2 3; do ncu -i /tmp/r2y_f.ncu-rep --page details --kernel-id :::$((k+1)) > $O/r2y_bits_f_details_k$k.txt 2>/dev/null; done
TSB_GRAPH=0 timeout 600 ncu --set full --clock-control none --cache-control none -k regex:k_merge_bits_f -s 7 -c 8 -o /tmp/r2y_w -f python tools/profile_plan.py --slices 4 > $O/r2y_ncu_w.log 2>&1
ncu -i /tmp/r2y_w.ncu-rep --page raw --csv > $O/r2y_bits_f_warm_full.csv 2>/dev/null
gzip -f $O/r2y_bench_launches.csv
