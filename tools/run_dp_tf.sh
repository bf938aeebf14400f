#!/bin/bash
N=${1:-2}
cd "$(dirname "$0")/.."
echo "== TF model test (refactored train_step)"; timeout 300 python -m pytest tests/test_gpu_tf_decoder.py -q -m gpu -k "tf_model" --tb=short 2>&1 | tail -4 | cut -c1-300
echo "== dp_check_tf N=$N"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 tests/manual/dp_check_tf.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -12 | cut -c1-300
