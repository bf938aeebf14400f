#!/bin/bash
mkdir -p gpurun_out
echo "== skinny timing"; timeout 120 python tools/skinny_timing.py 2>&1 | tail -4
echo "== tests"; timeout -k 10 400 python -m pytest tests -m gpu -q --timeout=120 -p no:cacheprovider --tb=short 2>&1 | tail -8 | cut -c1-300
for opts in "conv_mc=0" "conv_mc=1"; do
  echo "== bench $opts"
  LO_OPTS=$opts timeout -k 10 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_opt.log 2> gpurun_out/bench_opt.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_opt.log').read().strip().splitlines()[-1])
    a=d['roofline_all']
    print("  ms/step %.2f  img/s %.0f  e2e %.2f ms  att %.1f us (%.2f)  conv %.2f ms (%.2f)  dec %.2f ms" % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], a['attention']['us_per_launch'], a['attention']['frac'], a['conv']['ms'], a['conv']['frac'], a['phases']['decoder_fwd_bwd_ms']))
except Exception as e:
    print("  FAILED", e); print(open('gpurun_out/bench_opt.err').read()[-800:])
PY
done
