#!/bin/bash
# usage: tools/ncu_summary.sh report.ncu-rep  -> key metrics + per-region instruction counts + top stalls
rep=$1
ncu -i $rep --page raw --csv > /tmp/raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/raw.csv')))
hdr=rows[0]; units=rows[1]; r=rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'launch__registers_per_thread']
for i,h in enumerate(hdr):
    if h in want:
        print("%-90s %-10s %s" % (h, units[i], r[i]))
PY
ncu -i $rep --page source --csv --print-source sass > /tmp/src.csv 2>/dev/null
python $(dirname $0)/ncu_hot.py /tmp/src.csv 0 all > /tmp/hot_all.txt
awk 'NR>1 { split($0,a," inst "); split(a[2],b," "); n=$1; inst=b[1]; blk=int(n/50); s[blk]+=inst; } END { for (i=0;i<80;i++) if (s[i]>100000) printf "%4d-%4d %12d  (%.1f per worker-warp-tile@16384 tiles)\n", i*50, i*50+49, s[i], s[i]/262144.0 }' /tmp/hot_all.txt
python $(dirname $0)/ncu_hot.py /tmp/src.csv 0 top | head -${2:-25}
