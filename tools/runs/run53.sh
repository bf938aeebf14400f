#!/bin/bash
# run 53: verification pass as the driver does it: full GPU suite, smoke, default bench (both arms)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider --tb=short 2>&1 > gpurun_out/pytest53.log; tail -4 gpurun_out/pytest53.log | cut -c1-300
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
echo "== bench (default flags)"; t0=$(date +%s)
timeout 900 python bench.py 2>gpurun_out/bench53.err | tail -1 > gpurun_out/bench53.json; echo "wall=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench53.json').read())
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], d['roofline']['kernel'][:50], round(d['roofline']['frac'],3), d['roofline']['traffic'])
print(d['decode']); print(d['cpu_baseline']); print(d['clocks'])
PY
echo "== reference arm"; t0=$(date +%s)
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 2>gpurun_out/ref53.err | tail -1 > gpurun_out/ref53.json; echo "wall=$(( $(date +%s) - t0 ))s"; cut -c1-400 gpurun_out/ref53.json
