#!/bin/bash
# run 40 (round 2): whole GPU suite after the round-2 host/parity work (cfg2 goldens, TF dropout state, Philox dropout,
# diversity penalty, attention export), smoke, both bench arms
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider --tb=short -s 2>&1 > gpurun_out/pytest40.log; tail -30 gpurun_out/pytest40.log | cut -c1-400
grep -h "cfg2\|cfg1 bf16" gpurun_out/pytest40.log | cut -c1-200
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench40.err | tail -1 > gpurun_out/bench40.json
tail -3 gpurun_out/bench40.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench40.json').read())
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline_all']['phases'], d['gpu_launches'], d.get('cpu_baseline'))
for k,v in d['roofline_all'].items():
    if k!='phases': print(k, round(v['frac'],3), v.get('ms',v.get('us_per_launch')))
print('dominant', d['roofline']['kernel'][:40], d['roofline']['frac'])
PY
echo "== reference arm"
LO_REF_LINEARITY=1 timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/ref40.err | tail -1 > gpurun_out/ref40.json
cut -c1-600 gpurun_out/ref40.json
