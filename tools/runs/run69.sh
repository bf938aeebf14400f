#!/bin/bash
# run 69: ncu --set full of the tensor-core attention backward (3 launches inside a train step)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on --profile-from-start off -k regex:attention_bwd_mma_kernel -s 20 -c 3 -f -o gpurun_out/r2b_att_bwd_mma python tools/profile_step.py > gpurun_out/prof69.log 2>&1
tail -2 gpurun_out/prof69.log
ncu -i gpurun_out/r2b_att_bwd_mma.ncu-rep --page raw --csv > gpurun_out/r2b_att_bwd_mma_raw.csv 2>/dev/null
ncu -i gpurun_out/r2b_att_bwd_mma.ncu-rep --page details --csv 2>/dev/null > gpurun_out/r2b_att_bwd_mma_details.csv
python - <<'PY'
import csv
rows = list(csv.reader(l for l in open('gpurun_out/r2b_att_bwd_mma_raw.csv') if l.startswith('"')))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
for w in want:
    if w in hdr:
        j = hdr.index(w)
        print("%-90s %s %s" % (w, [r[j] for r in data], units[j]))
PY
