#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/sweep5.log
for s in 0 1; do
  MPET_PC_STREAMS=$s MPET_SPM_CFG=4 timeout 300 python scripts/sweep_pipe.py cfg5 72 pc >> gpurun_out/sweep5.log 2>&1 || echo "streams $s failed rc=$?" >> gpurun_out/sweep5.log
done
MPET_SPM_LANES=4 MPET_SPM_CFG=4 timeout 300 python scripts/sweep_pipe.py cfg5 72 pc >> gpurun_out/sweep5.log 2>&1 || echo "lanes 4 failed rc=$?" >> gpurun_out/sweep5.log
MPET_PC_STREAMS=0 MPET_SPM_CFG=4 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_spmm_pipeILi4ELi8ELi1E -s 2 -c 1 -f -o gpurun_out/full_spmm_p2 python scripts/sweep_pipe.py cfg5 72 pc > gpurun_out/ncu_full_spmm_p2.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
grep -E "pc_apply|failed" gpurun_out/sweep5.log
tail -3 gpurun_out/ncu_full_spmm_p2.log
