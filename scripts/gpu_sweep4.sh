#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/sweep4.log
for l in 8 16 32; do
  MPET_SPM_LANES=$l MPET_SPM_CFG=4 timeout 300 python scripts/sweep_pipe.py cfg5 72 pc >> gpurun_out/sweep4.log 2>&1 || echo "lanes $l failed rc=$?" >> gpurun_out/sweep4.log
done
MPET_SPM_CFG=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmm_pipe -s 0 -c 2 -f -o gpurun_out/full_spmm_p2 python scripts/sweep_pipe.py cfg5 72 pc > gpurun_out/ncu_full_spmm_p2.log 2>&1
grep -E "pc_apply|failed" gpurun_out/sweep4.log
