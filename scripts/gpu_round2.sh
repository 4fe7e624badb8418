#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --rtol 1e-2 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 300 python scripts/probe.py cfg5 72 1e-2 > gpurun_out/probe.log 2>&1; grep -E "assemble_lhs|spmv|rhs_prev" gpurun_out/probe.log
