#!/bin/bash
# run 41: fused decoder forward step (attention -> [gates GEMM + LSTM | grid barrier | next projection]) A/B, new tests
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest (new / touched)"; timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_decode.py tests/test_gpu_parity.py tests/test_gpu_tf_decoder.py -m gpu -q --timeout=600 -p no:cacheprovider --tb=short -s 2>&1 > gpurun_out/pytest41.log; tail -15 gpurun_out/pytest41.log | cut -c1-400
grep -h "cfg5 bf16" gpurun_out/pytest41.log | cut -c1-300
for f in 1 0; do
echo "== bench dec_fuse=$f"
LO_OPTS=dec_fuse=$f timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline --skip-decode 2>gpurun_out/bench41_$f.err | tail -1 > gpurun_out/bench41_$f.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench41_$f.json').read())
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline_all']['phases'], d['gpu_launches'])
PY
done
echo "== bench full (decode probe)"
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline 2>gpurun_out/bench41.err | tail -1 > gpurun_out/bench41.json
python -c "
import json
d=json.loads(open('gpurun_out/bench41.json').read()); print(d.get('decode'))"
tail -3 gpurun_out/bench41.err
