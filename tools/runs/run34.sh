#!/bin/bash
# run 34: vectorised column sums (bias gradients): whole suite + step time
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider --tb=short 2>&1 | tail -8 | cut -c1-300
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench"
timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline 2>gpurun_out/bench34.err | tail -1 > gpurun_out/bench34.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench34.json').read())
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline_all']['phases'])
for k,v in d['roofline_all'].items():
    if k!='phases': print(k, round(v['frac'],3), v.get('ms',v.get('us_per_launch')))
PY
