#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== TF decoder tests"
timeout 600 python -m pytest tests/test_gpu_tf_decoder.py -q -m gpu 2>&1 | tail -30
echo "== TF decoder bench"
timeout 600 python tools/tf_bench.py 2>&1 | tail -5
