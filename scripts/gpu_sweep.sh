#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep.log
for c in 0 1 2 3 4 5; do
  MPET_BLK_CFG=$c timeout 300 python scripts/sweep_pipe.py cfg5 72 spmv >> gpurun_out/sweep.log 2>&1 || echo "BLK $c failed rc=$?" >> gpurun_out/sweep.log
done
for c in 0 1 2 3 4 5; do
  MPET_SPM_CFG=$c timeout 300 python scripts/sweep_pipe.py cfg5 72 pc >> gpurun_out/sweep.log 2>&1 || echo "SPM $c failed rc=$?" >> gpurun_out/sweep.log
done
grep -E "spmv|pc_apply|failed" gpurun_out/sweep.log
