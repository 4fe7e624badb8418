#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/n2knobs.log
run() { echo "== $*" >> gpurun_out/n2knobs.log; env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 1 --warmup 3 2>gpurun_out/n2knobs.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['krylov_iterations'], round(d['ms_per_step']), round(d['value']), d['roofline']['share_of_step'])" >> gpurun_out/n2knobs.log; }
run MPET_P_CYCLES=3
run MPET_P_CYCLES=4
run MPET_P_CYCLES=4 MPET_P_DEGREE=6
cat gpurun_out/n2knobs.log
