#!/bin/bash
mkdir -p gpurun_out
echo "== diag"; timeout 300 python tests/manual/diag_traj.py tiny_eval > gpurun_out/diag.log 2>&1; tail -40 gpurun_out/diag.log
echo "== cpu probe"; timeout 240 python tests/manual/cpu_threads_probe.py > gpurun_out/cpu_probe.log 2>&1; tail -8 gpurun_out/cpu_probe.log
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider --tb=short --ignore=tests/test_gpu_tc.py 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
echo "== TC"; timeout -k 10 300 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout=120 -p no:cacheprovider --tb=short 2>&1 | tail -60 > gpurun_out/pytest_tc.log; tail -30 gpurun_out/pytest_tc.log
echo "== bench tc"; timeout -k 10 400 python bench.py --steps 5 --warmup 3 --kernels tc --skip-cpu-baseline > gpurun_out/bench_tc.log 2> gpurun_out/bench_tc.err; tail -2 gpurun_out/bench_tc.log; tail -12 gpurun_out/bench_tc.err
