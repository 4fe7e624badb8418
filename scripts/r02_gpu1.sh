#!/bin/bash
# round 2, call 1: GPU test suite + short bench on one B200
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02_gpu1_smi.txt 2>&1
nproc >> gpurun_out/r02_gpu1_smi.txt; free -g >> gpurun_out/r02_gpu1_smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_gpu1_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_gpu1_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-n 20 --no-configs > gpurun_out/r02_gpu1_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/r02_gpu1_bench.log
tail -5 gpurun_out/r02_gpu1_tests.log
tail -c 3000 gpurun_out/r02_gpu1_bench.log
