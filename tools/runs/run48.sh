#!/bin/bash
# run 48: round-2 profiles.  (1) launch list of the bench command (durations only); (2) ncu --set full of the step kernels
# inside one eager train step (in situ: --cache-control none keeps the step's cache state)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --skip-decode > gpurun_out/bench_under_ncu48.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_bench_launches.csv > gpurun_out/r2_bench_launches_summary.txt; head -14 gpurun_out/r2_bench_launches_summary.txt
gzip -f gpurun_out/r2_bench_launches.csv
prof() {   # name regex skip count
  timeout 900 ncu --set full --clock-control none --cache-control none --import-source on --profile-from-start off -k regex:$2 -s $3 -c $4 -f -o gpurun_out/r2_$1 python tools/profile_step.py > gpurun_out/prof48_$1.log 2>&1
  ncu -i gpurun_out/r2_$1.ncu-rep --page raw --csv > gpurun_out/r2_$1_raw.csv 2>/dev/null
  tail -1 gpurun_out/prof48_$1.log
}
prof att_fwd attention_fwd_pipe_kernel 70 3
prof att_bwd attention_bwd_mask_kernel 70 3
prof conv_p tc_conv_p_kernel 0 10
prof wgrad tc_wgrad_kernel 0 5
prof skinny skinny_mma_kernel 300 3
python tools/ncu_traffic.py gpurun_out/r2_att_fwd_raw.csv gpurun_out/r2_att_bwd_raw.csv gpurun_out/r2_conv_p_raw.csv gpurun_out/r2_wgrad_raw.csv gpurun_out/r2_skinny_raw.csv
cp profiles/r2_traffic.json gpurun_out/r2_traffic.json
