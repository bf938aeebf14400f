#!/bin/bash
mkdir -p gpurun_out
echo "== pytest with pdl=1"; LO_OPTS=pdl=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_decode.py -m gpu -q --timeout=300 -p no:cacheprovider --tb=short 2>&1 | tail -30 > gpurun_out/pytest_pdl.log; tail -12 gpurun_out/pytest_pdl.log | cut -c1-300
for opts in "pdl=0" "pdl=1"; do
  echo "== bench $opts"
  LO_OPTS=$opts timeout -k 10 300 python bench.py --steps 10 --warmup 3 --kernels tc --skip-cpu-baseline > gpurun_out/bench_opt.log 2> gpurun_out/bench_opt.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_opt.log').read().strip().splitlines()[-1])
    a=d['roofline_all']
    print("  ms/step %.2f  img/s %.0f  e2e %.2f ms  att %.1f us  conv %.2f ms  dec %.2f ms" % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], a['attention']['us_per_launch'], a['conv']['ms'], a['phases']['decoder_fwd_bwd_ms']))
except Exception as e:
    print("  FAILED", e); print(open('gpurun_out/bench_opt.err').read()[-800:])
PY
done
