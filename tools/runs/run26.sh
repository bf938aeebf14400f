#!/bin/bash
cd "$(dirname "$0")/../.."
echo "== new tests"
timeout 600 python -m pytest tests/test_gpu_tf_decoder.py tests/test_gpu_parity.py -q -m gpu -k "tf_ or cnn_variant" 2>&1 | tail -40 | cut -c1-400
