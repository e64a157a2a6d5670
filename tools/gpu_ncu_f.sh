#!/bin/bash
O=gpurun_out; mkdir -p $O
for p in 0 2; do
TSB_FUSE_PAIR=$p timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_merge_bits_f -c 3 -o /tmp/f_pair$p -f python tools/profile_plan.py --slices 1 > $O/f_pair$p.log 2>&1
ncu -i /tmp/f_pair$p.ncu-rep --page raw --csv > $O/f_pair$p.csv 2>/dev/null
ncu -i /tmp/f_pair$p.ncu-rep --page source --csv --kernel-id :::1 > $O/f_pair${p}_source.csv 2>/dev/null
ncu -i /tmp/f_pair$p.ncu-rep --page details --kernel-id :::1 > $O/f_pair${p}_details.txt 2>/dev/null
done
du -sh $O
