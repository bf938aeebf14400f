#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== cnn variant test"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k cnn_variant --tb=short 2>&1 | tail -8 | cut -c1-300
echo "== ncu full: attention_bwd_pipe_kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_bwd_pipe_kernel --launch-skip 60 -c 1 -o gpurun_out/att_bwd_pipe_r1 -f \
  python tools/profile_step.py > gpurun_out/prof_att_bwd.log 2>&1
ls -la gpurun_out/att_bwd_pipe_r1.ncu-rep
echo "== L2 window probe"
timeout 300 python tools/l2_window_probe.py 2>&1 | tail -14 | cut -c1-300
