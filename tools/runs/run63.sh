#!/bin/bash
# run 63: cluster-fused decoder step kernels (dec_cl, dec_cl_bwd): parity (optional schedules + golden fixtures) and A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_kernels.py -m gpu -q --timeout=600 -p no:cacheprovider --tb=short 2>&1 | tail -25 | cut -c1-600
for o in "dec_cl=1,dec_cl_bwd=1" "dec_cl=1,dec_cl_bwd=0" "dec_cl=0,dec_cl_bwd=1" "dec_cl=0,dec_cl_bwd=0" "att_policy_enc=2" "att_policy_enc=0"; do
echo "== bench $o"
LO_OPTS=$o timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu-baseline --skip-decode 2>gpurun_out/bench63.err | tail -1 > gpurun_out/bench63.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench63.json').read())
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {k: round(v,3) for k,v in d['roofline_all']['phases'].items() if not isinstance(v, dict)}, d['config']['loss_after'])
PY
tail -2 gpurun_out/bench63.err
done
