#!/bin/bash
# run 73: LSTM pointwise kernels fetch their loop-invariant operands before griddepcontrol.wait; two accumulator sets in the skinny GEMM:
# parity (full suite) + chain latency + bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider --tb=short -x 2>&1 | tail -8 | cut -c1-400
echo "== chain"; timeout 300 python tools/skinny_chain.py 2>&1 | grep -E "N 3072|N 2048"
echo "== breakdown"; timeout 600 python tools/dec_breakdown.py 2>&1 | tail -3
echo "== bench"
timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu-baseline --skip-decode 2>gpurun_out/bench73.err | tail -1 > gpurun_out/bench73.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench73.json').read())
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {k: round(v,3) for k,v in d['roofline_all']['phases'].items() if not isinstance(v, dict)}, d['config']['loss_after'])
PY
tail -1 gpurun_out/bench73.err
