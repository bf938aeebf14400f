#!/bin/bash
# run 23: TF-flavour decoder parity tests, then the whole GPU suite
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== TF decoder tests"
timeout 600 python -m pytest tests/test_gpu_tf_decoder.py -q -m gpu 2>&1 | tail -40
echo "== full suite"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
