#!/bin/bash
# run 75: launch list of one eager train step in the current state (per-launch times are cold-cache and serialised: use the shares)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/r2b_step_launches.csv python tools/profile_step.py > gpurun_out/prof75.log 2>&1
python tools/summarize_launches.py gpurun_out/r2b_step_launches.csv | tee gpurun_out/r2b_step_launches_summary.txt | head -48
