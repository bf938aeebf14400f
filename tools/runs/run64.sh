#!/bin/bash
# run 64: where the time of the cluster-fused step kernels goes (decoder loops with parts switched off), cluster on / off
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for o in "dec_cl=1,dec_cl_bwd=1" "dec_cl=0,dec_cl_bwd=0"; do
echo "== breakdown $o"; LO_OPTS=$o timeout 600 python tools/dec_breakdown.py 2>&1 | tail -9
done
