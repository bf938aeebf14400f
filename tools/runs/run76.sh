#!/bin/bash
# run 76: attention kernels draw 16-row chunks dynamically within a cluster (att_dynamic): parity + A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q --timeout=300 -p no:cacheprovider --tb=short -x 2>&1 | tail -12 | cut -c1-500
for o in "att_dynamic=1" "att_dynamic=0"; do
echo "== bench $o"
LO_OPTS=$o timeout 300 python bench.py --steps 30 --warmup 3 --skip-cpu-baseline --skip-decode 2>gpurun_out/bench76.err | tail -1 > gpurun_out/bench76.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench76.json').read())
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {k: round(v,3) for k,v in d['roofline_all']['phases'].items() if not isinstance(v, dict)}, d['config']['loss_after'])
PY
tail -2 gpurun_out/bench76.err
done
