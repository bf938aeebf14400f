#!/bin/bash
# run 65: per-CTA timeline of the cluster-fused step kernels (timing build)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python tools/cl_timeline.py 2>&1 | tail -60 | tee gpurun_out/cl_timeline65.txt
