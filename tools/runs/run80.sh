#!/bin/bash
# run 80: final default bench line (N=1) + ncu launch list of the bench command (round 2b)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python bench.py 2>gpurun_out/bench80.err | tail -1 > gpurun_out/bench80.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench80.json').read())
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], d['roofline']['kernel'][:40], round(d['roofline']['frac'],3))
dd=d['decode']; print({k: (round(v) if isinstance(v,float) and v>100 else v) for k,v in dd.items() if k!='workload' and k!='oracle_sample'})
print(d['cpu_baseline']['value'], d['clocks'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2b_bench_launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --skip-decode > gpurun_out/prof80.log 2>&1
python tools/summarize_launches.py gpurun_out/r2b_bench_launches.csv | tee gpurun_out/r2b_bench_launches_summary.txt | head -14
