#!/bin/bash
mkdir -p gpurun_out
echo "== skinny timing (<64,8> config)"; timeout 120 python tools/skinny_timing.py 2>&1 | tail -4 | cut -c1-420
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider --tb=short 2>&1 | tail -8 | cut -c1-300
echo "== decode bench"; timeout 600 python tools/decode_bench.py 2>&1 | tail -2 | cut -c1-500
