#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider --tb=short 2>&1 | tail -6 | cut -c1-300
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench full"; timeout -k 10 600 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; tail -4 gpurun_out/bench_full.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_full.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks','cpu_baseline')})
print(d['roofline']); print(d['roofline_all']['conv']); print(d['roofline_all']['phases'])
PY
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1c.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r1c.csv | head -32
