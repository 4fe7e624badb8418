#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -3 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?" >> gpurun_out/bench_n2.err
cat gpurun_out/bench_n2.json; tail -2 gpurun_out/bench_n2.err
MPET_NO_LANE1=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 1 --warmup 3 > gpurun_out/bench_n2_nolane.json 2>> gpurun_out/bench_n2.err
cat gpurun_out/bench_n2_nolane.json
