#!/bin/bash
# round 2: ncu capture WITHOUT cache flush between kernels (warm L2 as in a run) of the per-slice launches
O=gpurun_out; mkdir -p $O
export TSB_BITSF_TWO_SIDED=1
TSB_GRAPH=0 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_merge_bits_f -s 7 -c 8 -o /tmp/r2t_f -f python tools/profile_plan.py --slices 4 > $O/r2t_ncu.log 2>&1
ncu -i /tmp/r2t_f.ncu-rep --page raw --csv > $O/r2t_bits_f_warm.csv 2>/dev/null
for k in 0 1 2 3; do ncu -i /tmp/r2t_f.ncu-rep --page details --kernel-id :::$((k+1)) > $O/r2t_bits_f_warm_details_k$k.txt 2>/dev/null; done
tail -5 $O/r2t_ncu.log
