#!/bin/bash
mkdir -p gpurun_out
for opts in "att_nsplit=4" "att_nsplit=2" "att_nsplit=9"; do
  echo "== bench $opts"
  LO_OPTS=$opts timeout -k 10 300 python bench.py --steps 10 --warmup 3 --kernels tc --skip-cpu-baseline > gpurun_out/bench_opt.log 2> gpurun_out/bench_opt.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_opt.log').read().strip().splitlines()[-1])
    a=d['roofline_all']
    print("  ms/step %.2f  img/s %.0f  att %.1f us (%.0f GB/s, %.2f)  conv %.2f ms  dec %.2f ms" % (d['ms_per_step'], d['value'], a['attention']['us_per_launch'], a['attention']['achieved'], a['attention']['frac'], a['conv']['ms'], a['phases']['decoder_fwd_bwd_ms']))
except Exception as e:
    print("  FAILED", e); print(open('gpurun_out/bench_opt.err').read()[-600:])
PY
done
echo "== ncu full attention pipe (nsplit 9)"
LO_OPTS=att_nsplit=9 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_fwd_pipe -s 20 -c 2 -o gpurun_out/att_fwd_pipe_r1 -f python tools/profile_step.py > gpurun_out/prof_att.log 2>&1; tail -2 gpurun_out/prof_att.log
