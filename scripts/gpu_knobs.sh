#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/knobs.log
run() { echo "== $*" >> gpurun_out/knobs.log; env "$@" timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>gpurun_out/knobs.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['krylov_iterations'], round(d['ms_per_step']), d['roofline']['share_of_step'])" >> gpurun_out/knobs.log; }
run MPET_P_CYCLES=4
run MPET_P_CYCLES=5
run MPET_P_CYCLES=2 MPET_P_DEGREE=4
run MPET_P_CYCLES=3 MPET_P_DEGREE=3
run MPET_P_CYCLES=3 MPET_P_DEGREE=4
cat gpurun_out/knobs.log
