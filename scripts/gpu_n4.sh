#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?" >> gpurun_out/bench_n$N.err
cat gpurun_out/bench_n$N.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n$N.err | tail -5
