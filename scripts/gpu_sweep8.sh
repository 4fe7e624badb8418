#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/sweep8.log
for u in 16 8; do
  echo "LANES_U=$u LANES_P=32" >> gpurun_out/sweep8.log
  MPET_BLK_LANES_U=$u timeout 300 python scripts/sweep_pipe.py cfg5 72 spmv >> gpurun_out/sweep8.log 2>&1 || echo "failed rc=$?" >> gpurun_out/sweep8.log
done
grep -E "LANES|spmv|failed" gpurun_out/sweep8.log
