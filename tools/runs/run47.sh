#!/bin/bash
# run 47: forward attention with the context accumulation on mma.sync (att_mma): tests + A/B ; N=1 only
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider --tb=short 2>&1 > gpurun_out/pytest47.log; tail -8 gpurun_out/pytest47.log | cut -c1-400
for o in "att_mma=1" "att_mma=0"; do
echo "== bench $o"
LO_OPTS=$o timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline --skip-decode 2>gpurun_out/bench47.err | tail -1 > gpurun_out/bench47.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench47.json').read())
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {k: round(v,3) for k,v in d['roofline_all']['phases'].items() if not isinstance(v, dict)}, d['config']['loss_after'])
print({k: (round(v['frac'],3), round(v.get('ms', v.get('us_per_launch')),3)) for k,v in d['roofline_all'].items() if k!='phases'})
PY
done
