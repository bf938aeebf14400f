#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== cnn variant + TF tests"; timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tf_decoder.py -q -s -m gpu -k "cnn_variant or tf_" --tb=short 2>&1 | grep -v "^$" | tail -30 | cut -c1-250
echo "== bench"
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/bench29.err | tail -1 > gpurun_out/bench29.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench29.json').read())
print(d['ms_per_step'], d['value'], d['e2e']['value'])
for k,v in d['roofline_all'].items():
    if k!='phases': print(k, round(v['frac'],3), v.get('ms',v.get('us_per_launch')))
PY
