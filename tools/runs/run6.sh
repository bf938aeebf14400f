#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider --tb=short --ignore=tests/test_gpu_tc.py 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
echo "== TC"; timeout -k 10 300 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout=120 -p no:cacheprovider --tb=short 2>&1 | tail -60 > gpurun_out/pytest_tc.log; tail -5 gpurun_out/pytest_tc.log
for opts in "att_pipe=0" "att_pipe=1,att_policy_enc=0,att_policy_att1=0" "att_pipe=1,att_policy_enc=1,att_policy_att1=2" "att_pipe=1,att_policy_enc=2,att_policy_att1=1" "att_pipe=1,att_policy_enc=1,att_policy_att1=1" "att_pipe=1,att_policy_enc=1,att_policy_att1=2,att_nsplit=9" "att_pipe=1,att_policy_enc=1,att_policy_att1=2,att_nsplit=14"; do
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
