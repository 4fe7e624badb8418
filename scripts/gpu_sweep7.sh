#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/sweep7.log
for c in 0 1 2 3 4; do
  MPET_SPM_CFG=$c timeout 300 python scripts/sweep_pipe.py cfg5 72 pc >> gpurun_out/sweep7.log 2>&1 || echo "SPM $c failed rc=$?" >> gpurun_out/sweep7.log
done
MPET_SPM_CFG=2 MPET_SPM_LANES_HI=4 timeout 300 python scripts/sweep_pipe.py cfg5 72 pc >> gpurun_out/sweep7.log 2>&1
timeout 600 python bench.py --mesh-n 91 --steps 1 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1 n=91', d['config']['krylov_iterations'], d['ms_per_step'])" >> gpurun_out/sweep7.log
grep -E "pc_apply|failed|N=1" gpurun_out/sweep7.log
