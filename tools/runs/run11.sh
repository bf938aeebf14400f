#!/bin/bash
mkdir -p gpurun_out
python - <<PY
import torch
p=torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size/2**20, "MB; persisting max", getattr(p,'persisting_l2_cache_max_size',None))
PY
echo "== attention tests (cluster)"; timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_decode.py -m gpu -q --timeout=120 -p no:cacheprovider --tb=short 2>&1 | tail -8 | cut -c1-300
for opts in "att_cluster=0" "att_cluster=1" "att_cluster=1,att_policy_enc=2,att_policy_att1=2" "att_cluster=1,l2_persist_mb=64" "att_cluster=1,l2_persist_mb=64,att_policy_enc=2,att_policy_att1=2" "att_cluster=1,l2_persist_mb=96,att_policy_enc=1,att_policy_att1=2"; do
  echo "== bench $opts"
  LO_OPTS=$opts timeout -k 10 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_opt.log 2> gpurun_out/bench_opt.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_opt.log').read().strip().splitlines()[-1])
    a=d['roofline_all']
    print("  ms/step %.2f  img/s %.0f  e2e %.2f ms  att %.1f us (%.2f)  conv %.2f ms  dec %.2f ms" % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], a['attention']['us_per_launch'], a['attention']['frac'], a['conv']['ms'], a['phases']['decoder_fwd_bwd_ms']))
except Exception as e:
    print("  FAILED", e); print(open('gpurun_out/bench_opt.err').read()[-800:])
PY
done
echo "== ncu full skinny gemm"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_gemm_conv_kernel -s 30 -c 2 -o gpurun_out/skinny_r1 -f python tools/profile_step.py > gpurun_out/prof_sk.log 2>&1; tail -2 gpurun_out/prof_sk.log
