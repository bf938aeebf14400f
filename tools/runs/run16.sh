#!/bin/bash
mkdir -p gpurun_out
echo "== conv timing"; timeout 120 python tools/conv_timing.py 2>&1 | tail -10 | cut -c1-600
echo "== skinny timing"; timeout 120 python tools/skinny_timing.py 2>&1 | tail -2 | cut -c1-420
echo "== bench"; timeout -k 10 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_opt.log 2> gpurun_out/bench_opt.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_opt.log').read().strip().splitlines()[-1]); a=d['roofline_all']
print("  ms/step %.2f  img/s %.0f  att %.1f us  conv %.2f ms (%.2f)  dec %.2f ms" % (d['ms_per_step'], d['value'], a['attention']['us_per_launch'], a['conv']['ms'], a['conv']['frac'], a['phases']['decoder_fwd_bwd_ms']))
PY
echo "== ncu skinny sampling"
timeout 300 ncu --set full --sampling-interval 0 --clock-control none --import-source on -k regex:tc_gemm_conv_kernel -s 10 -c 4 -o gpurun_out/skinny_r1b -f python tools/skinny_loop.py > gpurun_out/prof_sk.log 2>&1; tail -1 gpurun_out/prof_sk.log
