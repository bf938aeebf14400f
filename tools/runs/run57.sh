#!/bin/bash
# run 57: extension (row-encoder biLSTM + second decoder layer): sequence-LSTM kernels and the whole model vs the CPU definition
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest ext"; timeout 900 python -m pytest tests/test_gpu_ext.py -m gpu -q --timeout=600 -p no:cacheprovider --tb=short 2>&1 > gpurun_out/pytest57.log; tail -40 gpurun_out/pytest57.log | cut -c1-300
