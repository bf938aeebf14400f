#!/bin/bash
# run 54: conv1 weight gradient from the forward's arg-max/ReLU codes: full suite + bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider --tb=short 2>&1 > gpurun_out/pytest54.log; tail -6 gpurun_out/pytest54.log | cut -c1-400
echo "== bench"
timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu-baseline --skip-decode 2>gpurun_out/bench54.err | tail -1 > gpurun_out/bench54.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench54.json').read())
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {k: round(v,3) for k,v in d['roofline_all']['phases'].items() if not isinstance(v, dict)}, d['config']['loss_after'])
print({k: (round(v['frac'],3), round(v.get('ms', v.get('us_per_launch')),3)) for k,v in d['roofline_all'].items() if k!='phases'})
PY
