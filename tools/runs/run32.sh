#!/bin/bash
# run 32: chunk-pipelined mma.sync GEMM: whole suite, step time, TF-flavour timings again
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider --tb=short 2>&1 | tail -8 | cut -c1-300
echo "== bench"
timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline 2>gpurun_out/bench32.err | tail -1 > gpurun_out/bench32.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench32.json').read())
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline_all']['phases'])
PY
echo "== TF bench"; timeout 300 python tools/tf_bench.py 2>&1 | tail -1 | cut -c1-600
