#!/bin/bash
# run 79: why the bench probe's beam-5 rate is below the device-level tool's
cd "$(dirname "$0")/../.."
timeout 600 python tools/decode_probe_dbg.py 2>&1 | tail -16
