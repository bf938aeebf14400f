#!/bin/bash
# run 74: attention kernels fetch the operands of their combine tail up front: parity + bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_decode.py -m gpu -q --timeout=600 -p no:cacheprovider --tb=short -x 2>&1 | tail -6 | cut -c1-400
echo "== bench"
timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu-baseline --skip-decode 2>gpurun_out/bench74.err | tail -1 > gpurun_out/bench74.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench74.json').read())
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {k: round(v,3) for k,v in d['roofline_all']['phases'].items() if not isinstance(v, dict)}, d['config']['loss_after'])
PY
tail -1 gpurun_out/bench74.err
