#!/bin/bash
# run 72: LSTM cell in the epilogue of the gates GEMM with bulk-copy operands and pre-wait epilogue operands (fuse_lstm): parity + A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; LO_OPTS=fuse_lstm=1 timeout 1200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_decode.py -m gpu -q --timeout=600 -p no:cacheprovider --tb=short 2>&1 | tail -8 | cut -c1-400
for o in "fuse_lstm=1" "fuse_lstm=0"; do
echo "== bench $o"
LO_OPTS=$o timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu-baseline --skip-decode 2>gpurun_out/bench72.err | tail -1 > gpurun_out/bench72.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench72.json').read())
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {k: round(v,3) for k,v in d['roofline_all']['phases'].items() if not isinstance(v, dict)}, d['config']['loss_after'])
PY
tail -1 gpurun_out/bench72.err
done
