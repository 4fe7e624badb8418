#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/sweep6.log
for h in 1 2 4; do
  echo "HI=$h LO=8" >> gpurun_out/sweep6.log
  MPET_SPM_LANES_HI=$h MPET_SPM_LANES_LO=8 timeout 300 python scripts/sweep_pipe.py cfg5 72 pc >> gpurun_out/sweep6.log 2>&1 || echo "failed rc=$?" >> gpurun_out/sweep6.log
done
for l in 1 2 4; do
  echo "HI=4 LO=$l" >> gpurun_out/sweep6.log
  MPET_SPM_LANES_HI=4 MPET_SPM_LANES_LO=$l timeout 300 python scripts/sweep_pipe.py cfg5 72 pc >> gpurun_out/sweep6.log 2>&1 || echo "failed rc=$?" >> gpurun_out/sweep6.log
done
grep -E "HI=|pc_apply|failed" gpurun_out/sweep6.log
