#!/bin/bash
O=gpurun_out; mkdir -p $O
TSB_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_merge_bits$' -c 40 -o /tmp/r2ad -f python tools/sampled_slices.py --config c4 --keep 6 --iters 1 > $O/r2ad_ncu.log 2>&1
ncu -i /tmp/r2ad.ncu-rep --page raw --csv > $O/r2ad_bits_general.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2ad_bits_general.csv')))
hdr=rows[0]; idx={h:i for i,h in enumerate(hdr)}
best=None
for n,r in enumerate(rows[2:]):
    t=float(r[idx['gpu__time_duration.sum']])
    if best is None or t>best[0]: best=(t,n,r)
t,n,r=best
print("slowest launch id", n+1, r[idx['Kernel Name']][:40], t)
for k in ['launch__grid_size','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','sm__cycles_active.avg','sm__cycles_elapsed.max','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']:
    if k in idx: print(k, r[idx[k]])
open('gpurun_out/r2ad_slowest_id.txt','w').write(str(n+1))
PY
ncu -i /tmp/r2ad.ncu-rep --page details --kernel-id :::$(cat $O/r2ad_slowest_id.txt) > $O/r2ad_details.txt 2>/dev/null
ncu -i /tmp/r2ad.ncu-rep --page source --csv --kernel-id :::$(cat $O/r2ad_slowest_id.txt) > $O/r2ad_source.csv 2>/dev/null
