#!/bin/bash
# run 78: beam-5 decode throughput regression hunt (cfg #5): per-step GEMM operand path on / off
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for o in "skinny_tma=1" "skinny_tma=0"; do
echo "== $o"; LO_OPTS=$o timeout 300 python tools/decode_bench.py 2>&1 | tail -1 | cut -c1-400
done
