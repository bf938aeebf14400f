#!/bin/bash
# run 27: cnn variant + fused-LSTM mma.sync epilogue + attention-bwd scalar prefetch: tests, then step time default vs fuse_lstm=1
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider --tb=short 2>&1 | tail -25 | cut -c1-300
echo "== bench default"
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/bench27a.err | tail -1 > gpurun_out/bench27a.json
echo "== bench fuse_lstm=1"
LO_OPTS="fuse_lstm=1" timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/bench27b.err | tail -1 > gpurun_out/bench27b.json
python - <<PY
import json
for f in ("a","b"):
    try:
        d=json.loads(open('gpurun_out/bench27%s.json'%f).read())
        print(f, d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline_all']['phases'], d['config']['loss_after'])
    except Exception as ex:
        print(f, 'failed', ex, open('gpurun_out/bench27%s.err'%f).read()[-800:])
PY
