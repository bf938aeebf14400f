#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider --tb=short --ignore=tests/test_gpu_tc.py 2>&1 | tail -70 > gpurun_out/pytest_gpu.log; tail -30 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== TC"; timeout -k 10 300 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout=120 -p no:cacheprovider --tb=short 2>&1 | tail -60 > gpurun_out/pytest_tc.log; tail -5 gpurun_out/pytest_tc.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench full"; timeout -k 10 600 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; tail -1 gpurun_out/bench_full.log | cut -c1-300; tail -8 gpurun_out/bench_full.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_full.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks','cpu_baseline')})
print(d['roofline']); print(d['roofline_all']['attention']); print(d['roofline_all']['phases'])
PY
echo "== bench reference arm"; timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | cut -c1-600
