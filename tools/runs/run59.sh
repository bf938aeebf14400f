#!/bin/bash
# run 59: two half-batch decoder chains on two streams (dec_streams=2) A/B with the round-2 kernels (mma.sync per-step GEMMs, mask bits)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for o in "dec_streams=1" "dec_streams=2" "dec_streams=2,att_nsplit=8"; do
echo "== bench $o"
LO_OPTS=$o timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu-baseline --skip-decode 2>gpurun_out/bench59.err | tail -1 > gpurun_out/bench59.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench59.json').read())
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {k: round(v,3) for k,v in d['roofline_all']['phases'].items() if not isinstance(v, dict)}, d['config']['loss_after'])
PY
tail -2 gpurun_out/bench59.err
done
