#!/bin/bash
O=gpurun_out; mkdir -p $O
TSB_GRAPH=0 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k 'regex:^k_merge_bits$' -s 5 -c 5 -o /tmp/r2am -f python tools/profile_plan.py --slices 1 > $O/r2am_ncu.log 2>&1
ncu -i /tmp/r2am.ncu-rep --page raw --csv > $O/r2am_expand.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2am_expand.csv')))
hdr=rows[0]; idx={h:i for i,h in enumerate(hdr)}
for k in ['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__registers_per_thread','sm__cycles_active.avg','sm__cycles_elapsed.max','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_drain_per_issue_active.ratio']:
    if k in idx: print(k.replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio',''), [r[idx[k]][:30] for r in rows[2:]])
PY
ncu -i /tmp/r2am.ncu-rep --page source --csv --kernel-id :::1 > $O/r2am_source.csv 2>/dev/null
