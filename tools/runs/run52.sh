#!/bin/bash
# run 52 (N GPUs): the driver's own command line at N GPUs — whole step incl. NCCL in one CUDA graph (default) — must print its
# line and EXIT (run 46 hung in ncclCommDestroy until the timeout)
N=${1:-8}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 2> gpurun_out/b52.err | tail -1 > gpurun_out/b52.json
echo "rc=$? wall=$(( $(date +%s) - t0 ))s"
python -c "
import json; d=json.loads(open('gpurun_out/b52.json').read()); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['config']['cuda_graph'], d['clocks'])" || tail -20 gpurun_out/b52.err
t0=$(date +%s)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2> gpurun_out/b52r.err | tail -1 | cut -c1-300
echo "reference arm rc=$? wall=$(( $(date +%s) - t0 ))s"
