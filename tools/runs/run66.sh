#!/bin/bash
# run 66: dependent chain of skinny GEMMs: latency vs rows read and vs number of CTAs
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python tools/skinny_chain.py 2>&1 | tail -20 | tee gpurun_out/skinny_chain66.txt
