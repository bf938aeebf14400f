#!/bin/bash
# run 21: skinny mma kernel tests + step time with/without it
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest (tc + kernels + decode)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench skinny_mma=1 (default)"
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/bench21a.err | tail -1 | tee gpurun_out/bench21a.json
echo
echo "== bench skinny_mma=0"
LO_OPTS="skinny_mma=0" timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/bench21b.err | tail -1 | tee gpurun_out/bench21b.json
