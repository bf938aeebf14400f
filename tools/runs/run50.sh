#!/bin/bash
# run 50: conv weight gradient with 128 x 256 tiles (wgrad256) A/B + full suite + launch list of one eager step
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider --tb=short 2>&1 > gpurun_out/pytest50.log; tail -6 gpurun_out/pytest50.log | cut -c1-400
for o in "wgrad256=1" "wgrad256=0"; do
echo "== bench $o"
LO_OPTS=$o timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline --skip-decode 2>gpurun_out/bench50.err | tail -1 > gpurun_out/bench50.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench50.json').read())
print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {k: round(v,3) for k,v in d['roofline_all']['phases'].items() if not isinstance(v, dict)}, d['config']['loss_after'])
print({k: (round(v['frac'],3), round(v.get('ms', v.get('us_per_launch')),3)) for k,v in d['roofline_all'].items() if k!='phases'})
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/r2_step_launches_v2.csv python tools/profile_step.py > gpurun_out/prof50.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_step_launches_v2.csv | tee gpurun_out/r2_step_launches_v2_summary.txt | head -45
