#!/bin/bash
# run 49: full GPU suite on the current build (batched tcgen05 alpha^T dctx, embedding-path GEMMs, TF optimisers, datt1 d w_full
# algebra) + bench; then the attention tuning variant (1 row per warp per stage, 3 CTAs/SM, 6-CTA clusters) A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider --tb=short 2>&1 > gpurun_out/pytest49.log; tail -8 gpurun_out/pytest49.log | cut -c1-400
for v in _C _C_rpw1; do
echo "== bench lib $v"
LO_LIB_DIR=$v timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline --skip-decode 2>gpurun_out/bench49.err | tail -1 > gpurun_out/bench49_$v.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench49_$v.json').read())
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {k: round(v,3) for k,v in d['roofline_all']['phases'].items() if not isinstance(v, dict)}, d['config']['loss_after'])
print({k: (round(v['frac'],3), round(v.get('ms', v.get('us_per_launch')),3)) for k,v in d['roofline_all'].items() if k!='phases'})
PY
tail -2 gpurun_out/bench49.err
done
echo "== pytest attention kernels on the variant"; LO_LIB_DIR=_C_rpw1 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q --timeout=600 -p no:cacheprovider --tb=short 2>&1 | tail -4 | cut -c1-300
