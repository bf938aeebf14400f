#!/bin/bash
# run 68: attention backward with alpha / d reg staged in shared memory before the wait: parity + bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q --timeout=600 -p no:cacheprovider --tb=short -x 2>&1 | tail -8 | cut -c1-400
for o in "att_bwd_mma=1" "att_bwd_mma=0"; do
echo "== bench $o"
LO_OPTS=$o timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu-baseline --skip-decode 2>gpurun_out/bench68.err | tail -1 > gpurun_out/bench68.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench68.json').read())
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {k: round(v,3) for k,v in d['roofline_all']['phases'].items() if not isinstance(v, dict)}, d['config']['loss_after'])
PY
tail -1 gpurun_out/bench68.err
done
