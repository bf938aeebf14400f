#!/bin/bash
# run 25: verification pass — GPU suite, smoke, full bench (both arms), launch list of the bench command, one full capture of the
# mma.sync per-step GEMM
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider --tb=short 2>&1 | tail -8 | cut -c1-300
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench --impl reference"; timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref25.json 2> gpurun_out/bench_ref25.err; tail -1 gpurun_out/bench_ref25.json | cut -c1-600
echo "== bench full"; timeout -k 10 600 python bench.py > gpurun_out/bench25.json 2> gpurun_out/bench25.err; tail -3 gpurun_out/bench25.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench25.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks','cpu_baseline')})
print(d['roofline']); print(d['roofline_all']['conv']); print(d['roofline_all']['conv_tensor_kernels']); print(d['roofline_all']['phases'])
PY
echo "== launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/launches_r1_bench.csv \
  python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r1_bench.csv | head -40
echo "== ncu full: skinny_mma_kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:skinny_mma_kernel --launch-skip 300 -c 2 -o gpurun_out/skinny_mma_r1 -f \
  python tools/profile_step.py > gpurun_out/prof_skinny_mma.log 2>&1
ls -la gpurun_out/skinny_mma_r1.ncu-rep
