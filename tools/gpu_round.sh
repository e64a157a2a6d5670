#!/bin/bash
# One GPU-box visit: parity tests, bench (packed and byte-per-entry), per-op timing, ncu launch list and full captures.
# Everything lands under gpurun_out/.  Usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r1b}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 700 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
timeout 300 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; tail -c 600 $O/${TAG}_bench_n1.json
timeout 200 python bench.py --no-pack --no-cpu-baseline > $O/${TAG}_bench_n1_nopack.json 2> $O/${TAG}_bench_n1_nopack.err
timeout 200 python tools/profile_plan.py --top 40 > $O/${TAG}_profile_plan_sat.txt 2>&1
timeout 200 python tools/profile_plan.py --top 40 --count > $O/${TAG}_profile_plan_count.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_merge_bits --launch-skip 60 -c 24 -o $O/${TAG}_merge_bits_full -f \
    python tools/profile_plan.py --slices 2 > $O/${TAG}_ncu_full.log 2>&1
ncu -i $O/${TAG}_merge_bits_full.ncu-rep --page raw --csv > $O/${TAG}_merge_bits_full.csv 2>/dev/null
ls -la $O | tail -20
