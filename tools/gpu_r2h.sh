#!/bin/bash
# round 2 evidence: ncu launch list of the bench command, full captures of the dominant kernel, a permute and a GEMM
O=gpurun_out; mkdir -p $O
TSB_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2_bench_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2_ncu_bench.log 2>&1
python tools/ncu_summary.py launches $O/r2_bench_launches.csv > $O/r2_bench_launches_summary.txt; head -12 $O/r2_bench_launches_summary.txt
gzip -f $O/r2_bench_launches.csv
TSB_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_merge_bits_f -s 192 -c 8 -f -o /tmp/r2_bits_f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-micro --c4-keep -1 > $O/r2_ncu_bits_f.log 2>&1
ncu -i /tmp/r2_bits_f.ncu-rep --page raw --csv > $O/r2_bits_f_full.csv 2>/dev/null
ncu -i /tmp/r2_bits_f.ncu-rep --page source --csv --kernel-id :::1 > $O/r2_bits_f_source_k1.csv 2>/dev/null
ncu -i /tmp/r2_bits_f.ncu-rep --page details --kernel-id :::1 > $O/r2_bits_f_details_k1.txt 2>/dev/null
ncu -i /tmp/r2_bits_f.ncu-rep --page details --kernel-id :::2 > $O/r2_bits_f_details_k2.txt 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:k_permute -s 4 -c 3 -f -o /tmp/r2_permute python tools/microbench.py permute --bits 30 --random 2 --iters 1 > $O/r2_ncu_permute.log 2>&1
ncu -i /tmp/r2_permute.ncu-rep --page raw --csv > $O/r2_permute_full.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:k_gemm_i8 -s 1 -c 2 -f -o /tmp/r2_gemm python tools/microbench.py gemm --iters 1 > $O/r2_ncu_gemm.log 2>&1
ncu -i /tmp/r2_gemm.ncu-rep --page raw --csv > $O/r2_gemm_full.csv 2>/dev/null
ls -la $O/r2_*; du -sh $O
