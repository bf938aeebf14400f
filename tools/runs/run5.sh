#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider --tb=short --ignore=tests/test_gpu_tc.py 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
echo "== TC"; timeout -k 10 300 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout=120 -p no:cacheprovider --tb=short 2>&1 | tail -60 > gpurun_out/pytest_tc.log; tail -30 gpurun_out/pytest_tc.log
echo "== bench tc"; timeout -k 10 400 python bench.py --steps 10 --warmup 3 --kernels tc --skip-cpu-baseline > gpurun_out/bench_tc.log 2> gpurun_out/bench_tc.err; tail -2 gpurun_out/bench_tc.log | cut -c1-400; tail -6 gpurun_out/bench_tc.err
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1
tail -2 gpurun_out/prof_step.log; python tools/summarize_launches.py gpurun_out/launches_r1.csv | head -40
