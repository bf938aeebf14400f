#!/bin/bash
# run 31: verification pass after the last kernel changes (mma.sync row blocks, beam logits): suite, smoke, decode throughput, both bench arms
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider --tb=short 2>&1 | tail -12 | cut -c1-300
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
echo "== decode bench"; timeout 300 python tools/decode_bench.py 2>&1 | tail -1 | cut -c1-600
echo "== TF bench"; timeout 300 python tools/tf_bench.py 2>&1 | tail -1 | cut -c1-600
echo "== bench --impl reference"; timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref31.json 2> gpurun_out/bench_ref31.err; tail -1 gpurun_out/bench_ref31.json | cut -c1-300
echo "== bench full"; timeout -k 10 600 python bench.py > gpurun_out/bench31.json 2> gpurun_out/bench31.err; tail -3 gpurun_out/bench31.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench31.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks','cpu_baseline')})
for k,v in d['roofline_all'].items():
    if k!='phases': print(k, round(v['frac'],3), v.get('ms',v.get('us_per_launch')))
print(d['roofline_all']['phases'])
PY
