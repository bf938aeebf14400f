#!/bin/bash
# run 62: per-CTA %globaltimer timeline of the attention step kernels (timing build _C_timing, -DLO_ATT_TIMING)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python tools/att_timeline.py 2>&1 | tail -45 | tee gpurun_out/att_timeline62.txt
