#!/bin/bash
# run 58: full GPU suite (incl. the extension) + the extension workload `bench.py --workload cfg4` at N=1
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider --tb=short 2>&1 > gpurun_out/pytest58.log; tail -5 gpurun_out/pytest58.log | cut -c1-300
echo "== bench cfg4"
timeout 600 python bench.py --workload cfg4 --steps 20 --warmup 3 2>gpurun_out/bench58.err | tail -1 > gpurun_out/bench58.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench58.json').read())
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], d['config']['loss_after'], d['config']['cuda_graph'])
PY
tail -3 gpurun_out/bench58.err
