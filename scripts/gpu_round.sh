#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list and full captures of the top kernels.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --rtol 1e-2 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
for k in k_block_rows_staged k_asm22 k_spmm_staged; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/full_$k \
    python bench.py --steps 1 --warmup 3 --rtol 1e-2 --no-cpu-baseline > gpurun_out/ncu_full_$k.log 2>&1
done
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err
