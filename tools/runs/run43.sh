#!/bin/bash
# run 43: launch list of one eager train step (ncu, durations only) to rank the hoisted decoder kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/r2_step_launches_v1.csv python tools/profile_step.py > gpurun_out/prof43.log 2>&1
tail -2 gpurun_out/prof43.log
python tools/summarize_launches.py gpurun_out/r2_step_launches_v1.csv | tee gpurun_out/r2_step_launches_v1_summary.txt | head -60
