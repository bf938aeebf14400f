#!/bin/bash
# N-GPU data-parallel check: gradient equivalence (tests/manual/dp_check.py), then the bench line at N ranks (eager, the default for N > 1)
N=${1:-4}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== dp_check N=$N"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tests/manual/dp_check.py 2>&1 | tail -4 | cut -c1-300
echo "== bench N=$N (eager)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 10 --warmup 3 --skip-cpu-baseline 2> gpurun_out/b${N}.err | tail -1 > gpurun_out/b${N}.json
python -c "import json; d=json.loads(open('gpurun_out/b${N}.json').read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['cuda_graph'], d['clocks'])" || tail -20 gpurun_out/b${N}.err
