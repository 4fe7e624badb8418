#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/iters.log
run1() { timeout 900 python bench.py --mesh-n $1 --steps 1 --warmup 3 --no-cpu-baseline 2>>gpurun_out/iters.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1 n=$1', d['config']['krylov_iterations'], d['ms_per_step'], d['config']['total_dofs'])" >> gpurun_out/iters.log; }
run2() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mesh-n $1 --steps 1 --warmup 3 2>>gpurun_out/iters.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=2 n=$1', d['config']['krylov_iterations'], d['ms_per_step'], d['config']['total_dofs'])" >> gpurun_out/iters.log; }
run2 24    # global 30^3
run1 30
run2 46    # global 58^3
run1 58
run1 91    # the global problem of the default N=2 bench
cat gpurun_out/iters.log
