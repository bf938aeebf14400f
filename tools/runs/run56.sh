#!/bin/bash
# run 56: ncu --set full of the ten tc_conv_p_kernel launches of one train step with the final kernels (MT=2 for <= 128 channels)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on --profile-from-start off -k regex:tc_conv_p_kernel -s 0 -c 10 -f -o gpurun_out/r2_conv_p python tools/profile_step.py > gpurun_out/prof56.log 2>&1
ncu -i gpurun_out/r2_conv_p.ncu-rep --page raw --csv > gpurun_out/r2_conv_p_raw.csv 2>/dev/null
tail -1 gpurun_out/prof56.log
python tools/ncu_traffic.py profiles/r2_att_fwd_raw.csv profiles/r2_att_bwd_raw.csv gpurun_out/r2_conv_p_raw.csv profiles/r2_wgrad_raw.csv profiles/r2_skinny_raw.csv
cp profiles/r2_traffic.json gpurun_out/r2_traffic.json
