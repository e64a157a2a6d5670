#!/bin/bash
# launch list of the bench under ncu, sampled-slice runs of the wide configs, 2-GPU run when 2 GPUs are visible
O=gpurun_out; mkdir -p $O
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r1c_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r1c_bench_under_ncu.log 2>&1
timeout 300 python tools/sampled_slices.py --config c4 > $O/r1c_sampled_c4.json 2> $O/r1c_sampled_c4.err
timeout 300 python tools/sampled_slices.py --config c4 --profile > $O/r1c_sampled_c4_per_op.txt 2>&1
timeout 400 python tools/sampled_slices.py --config c5 > $O/r1c_sampled_c5.json 2> $O/r1c_sampled_c5.err
timeout 300 python tools/sampled_slices.py --config c2 > $O/r1c_sampled_c2.json 2> $O/r1c_sampled_c2.err
tail -c 700 $O/r1c_sampled_c4.json; echo; tail -c 700 $O/r1c_sampled_c5.json; echo; tail -c 500 $O/r1c_sampled_c2.json
