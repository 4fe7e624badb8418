#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/sweep3.log
nproc >> gpurun_out/sweep3.log; python -c "import os; print('cpus', os.cpu_count(), len(os.sched_getaffinity(0)))" >> gpurun_out/sweep3.log
for o in 0 1; do
  MPET_BLK_ORDER=$o timeout 300 python scripts/sweep_pipe.py cfg5 72 spmv >> gpurun_out/sweep3.log 2>&1 || echo "BLK order $o failed rc=$?" >> gpurun_out/sweep3.log
done
for c in 4 6 7; do
  MPET_SPM_ORDER=1 MPET_SPM_CFG=$c timeout 300 python scripts/sweep_pipe.py cfg5 72 pc >> gpurun_out/sweep3.log 2>&1 || echo "SPM $c failed rc=$?" >> gpurun_out/sweep3.log
done
grep -E "cpus|spmv|pc_apply|failed" gpurun_out/sweep3.log
