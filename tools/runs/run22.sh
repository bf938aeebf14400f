#!/bin/bash
# run 22: PDL pre-wait prefetch (attention producer, skinny weights) — tests + step time + tc-only conv probe
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== bench"
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/bench22.err | tail -1 | tee gpurun_out/bench22.json
