#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
echo "== option tests"; timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -k optional --timeout=120 -p no:cacheprovider --tb=short 2>&1 | tail -4 | cut -c1-300
echo "== bench N=$N eager"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --skip-cpu-baseline 2> gpurun_out/bN.err | tail -1 > gpurun_out/bN.log; python -c "import json; d=json.loads(open('gpurun_out/bN.log').read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['cuda_graph'])" || tail -20 gpurun_out/bN.err
echo "== bench N=$N graph"; LO_DP_GRAPH=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --skip-cpu-baseline 2> gpurun_out/bNg.err | tail -1 > gpurun_out/bNg.log; python -c "import json; d=json.loads(open('gpurun_out/bNg.log').read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['cuda_graph'])" || tail -30 gpurun_out/bNg.err
