#!/bin/bash
# run 55: conv layers with <= 128 output channels: two 128-position sub-tiles per CTA tile sharing each weight stage (conv_mt2) A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_kernels.py -m gpu -q --timeout=600 -p no:cacheprovider --tb=short 2>&1 | tail -5 | cut -c1-400
for o in "conv_mt2=1" "conv_mt2=0"; do
echo "== bench $o"
LO_OPTS=$o timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu-baseline --skip-decode 2>gpurun_out/bench55.err | tail -1 > gpurun_out/bench55.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench55.json').read())
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), {k: round(v,3) for k,v in d['roofline_all']['phases'].items() if not isinstance(v, dict)}, d['config']['loss_after'])
print({k: (round(v['frac'],3), round(v.get('ms', v.get('us_per_launch')),3)) for k,v in d['roofline_all'].items() if k!='phases'})
PY
done
