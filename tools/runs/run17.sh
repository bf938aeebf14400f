#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout -k 10 400 python -m pytest tests -m gpu -q --timeout=120 -p no:cacheprovider --tb=short 2>&1 | tail -6 | cut -c1-300
echo "== ncu skinny sampling (20 launches)"
timeout 300 ncu --section SourceCounters --section WarpStateStats --clock-control none --import-source on -k regex:tc_gemm_conv_kernel -c 20 -o gpurun_out/skinny_r1b -f python tools/skinny_loop.py > gpurun_out/prof_sk.log 2>&1; tail -1 gpurun_out/prof_sk.log
