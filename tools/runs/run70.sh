#!/bin/bash
# run 70: per-CTA timeline of the grid-barrier fused forward step kernel (timing build)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python tools/fuse_timeline.py 2>&1 | tail -14 | tee gpurun_out/fuse_timeline70.txt
