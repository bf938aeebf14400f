#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
echo "== dp_check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/manual/dp_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -8
echo "== bench N=1"; timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline 2> gpurun_out/b1.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
echo "== bench N=1 no graph"; timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --no-graph 2> gpurun_out/b1.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --skip-cpu-baseline 2> gpurun_out/bN.err | tail -1 > gpurun_out/bN.log; python -c "import json; d=json.loads(open('gpurun_out/bN.log').read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['cuda_graph'])" || tail -20 gpurun_out/bN.err
