#!/bin/bash
# run 33: attention kernels fetch forward-saved operands before griddepcontrol.wait: suite, step time, fresh full captures
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider --tb=short 2>&1 | tail -8 | cut -c1-300
echo "== bench"
timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline 2>gpurun_out/bench33.err | tail -1 > gpurun_out/bench33.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench33.json').read())
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline_all']['phases'])
PY
echo "== ncu full: attention fwd + bwd (one launch each, in situ)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_.*_pipe_kernel --launch-skip 120 -c 2 -o gpurun_out/att_pipe_r1_final -f \
  python tools/profile_step.py > gpurun_out/prof_att_final.log 2>&1
ls -la gpurun_out/att_pipe_r1_final.ncu-rep
