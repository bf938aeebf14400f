#!/bin/bash
# run 77: verification pass as the driver does it (round 2b): full GPU suite, smoke, default bench (both arms)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider --tb=short 2>&1 > gpurun_out/pytest77.log; tail -4 gpurun_out/pytest77.log | cut -c1-300
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
echo "== bench (default flags)"; t0=$(date +%s)
timeout 900 python bench.py 2>gpurun_out/bench77.err | tail -1 > gpurun_out/bench77.json; echo "wall=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench77.json').read())
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], d['roofline']['kernel'][:50], round(d['roofline']['frac'],3), d['roofline']['traffic'])
print(d['decode']); print(d['cpu_baseline']); print(d['clocks'])
PY
echo "== reference arm"; t0=$(date +%s)
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 2>gpurun_out/ref77.err | tail -1 > gpurun_out/ref77.json; echo "wall=$(( $(date +%s) - t0 ))s"; cut -c1-400 gpurun_out/ref77.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench77.json').read())
print({k: (round(v['frac'],3), round(v.get('ms', v.get('us_per_launch')),3)) for k,v in d['roofline_all'].items() if k!='phases'})
print(d['roofline_all']['phases'])
PY
