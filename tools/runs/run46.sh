#!/bin/bash
# run 46 (N GPUs): data-parallel parity (per-layer encoder buckets) + bench with the step incl. NCCL in one CUDA graph (default) vs eager
N=${1:-2}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== dp_check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/manual/dp_check.py 2>&1 | grep -v "^W\|^\[W\|warn" | tail -6
for g in 1 0; do
echo "== bench N=$N LO_DP_GRAPH=$g"
LO_DP_GRAPH=$g timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$g bench.py --gpus $N --steps 30 --warmup 5 2> gpurun_out/b46_$g.err | tail -1 > gpurun_out/b46_$g.json
python -c "
import json; d=json.loads(open('gpurun_out/b46_$g.json').read()); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['config']['cuda_graph'])" || tail -20 gpurun_out/b46_$g.err
done
echo "== bench N=1 same box"
timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline --skip-decode 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"
