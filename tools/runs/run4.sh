#!/bin/bash
mkdir -p gpurun_out
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1
tail -2 gpurun_out/prof_step.log; python tools/summarize_launches.py gpurun_out/launches_r1.csv | head -45
echo "== ncu full attention fwd"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_fwd_kernel -s 20 -c 2 -o gpurun_out/att_fwd_r1 -f python tools/profile_step.py > gpurun_out/prof_att.log 2>&1; tail -2 gpurun_out/prof_att.log
echo "== ncu full tc conv"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_gemm_conv_kernel -s 3 -c 3 -o gpurun_out/tc_conv_r1 -f python tools/profile_step.py > gpurun_out/prof_conv.log 2>&1; tail -2 gpurun_out/prof_conv.log
ls -la gpurun_out/*.ncu-rep
