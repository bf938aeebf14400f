#!/bin/bash
# run 67: fused step kernels with the spread-counter grid barrier (dec_fuse, dec_fuse_bwd): parity + A/B + breakdown
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout=600 -p no:cacheprovider --tb=short -k "optional" 2>&1 | tail -8 | cut -c1-400
for o in "dec_fuse=1,dec_fuse_bwd=1" "dec_fuse=1" "dec_fuse_bwd=1" "dec_fuse=0"; do
echo "== bench $o"
LO_OPTS=$o timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu-baseline --skip-decode 2>gpurun_out/bench67.err | tail -1 > gpurun_out/bench67.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench67.json').read())
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {k: round(v,3) for k,v in d['roofline_all']['phases'].items() if not isinstance(v, dict)}, d['config']['loss_after'])
PY
tail -1 gpurun_out/bench67.err
done
