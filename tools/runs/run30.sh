#!/bin/bash
# run 30: attention kernels with separate att1/enc widths: whole GPU suite, TF-flavour timings, decode throughput, step time
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider --tb=short 2>&1 | tail -12 | cut -c1-300
echo "== TF bench"; timeout 300 python tools/tf_bench.py 2>&1 | tail -2 | cut -c1-600
echo "== decode bench"; timeout 300 python tools/decode_bench.py 2>&1 | tail -2 | cut -c1-600
echo "== bench"
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/bench30.err | tail -1 > gpurun_out/bench30.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench30.json').read())
print(d['ms_per_step'], d['value'], d['e2e']['value'])
for k,v in d['roofline_all'].items():
    if k!='phases': print(k, round(v['frac'],3), v.get('ms',v.get('us_per_launch')))
PY
