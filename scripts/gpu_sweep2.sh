#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/sweep2.log
for c in 4 6 7 8; do
  MPET_SPM_CFG=$c timeout 300 python scripts/sweep_pipe.py cfg5 72 pc >> gpurun_out/sweep2.log 2>&1 || echo "SPM $c failed rc=$?" >> gpurun_out/sweep2.log
done
for c in 4 6; do
MPET_SPM_CFG=$c timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_spmm_pipe -c 120 --csv --log-file gpurun_out/pc_launches_$c.csv python scripts/sweep_pipe.py cfg5 72 pc > gpurun_out/ncu_pc_$c.log 2>&1
done
grep -E "pc_apply|failed" gpurun_out/sweep2.log
