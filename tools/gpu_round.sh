#!/bin/bash
# One GPU-box visit: parity tests, bench (packed and byte-per-entry), reference arm, per-op timing, ncu launch list and
# a full capture of the dominant kernel.  Everything lands under gpurun_out/ (reports are converted to CSV there; the
# .ncu-rep files stay on the box).  Usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r1e}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 700 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
timeout 120 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 300 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; tail -c 400 $O/${TAG}_bench_n1.json; echo
timeout 200 python bench.py --no-pack --no-cpu-baseline > $O/${TAG}_bench_n1_nopack.json 2> $O/${TAG}_bench_n1_nopack.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_reference_arm.err
timeout 200 python tools/profile_plan.py --top 30 > $O/${TAG}_profile_plan_sat.txt 2>&1
timeout 200 python tools/profile_plan.py --top 30 --count > $O/${TAG}_profile_plan_count.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${TAG}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_merge_bits_f --launch-skip 8 -c 8 -o /tmp/${TAG}_bits_f_full -f \
    python tools/profile_plan.py --slices 3 > $O/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_bits_f_full.ncu-rep --page raw --csv > $O/${TAG}_bits_f_full.csv 2>/dev/null
du -sh $O
