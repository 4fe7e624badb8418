#!/bin/bash
# round 2: P1-field cycle knobs under the PETSc-default tolerance (1 GPU, cfg5)
mkdir -p gpurun_out
run() {  # name, env
  env $2 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_sw1_$1.log 2>&1
  python - <<P
import json
for line in open('gpurun_out/r02_sw1_$1.log'):
    if line.startswith('{"metric"'):
        d=json.loads(line); r=d["roofline"]
        print("$1", round(d["ms_per_step"],1), d["config"]["krylov_iterations"], "ms/it", round(r["ms_per_iteration"],3), "pc", round(r["preconditioner"]["avg_application_ms"],3), "launches/it", round(r["launches_per_iteration"]))
P
}
run base "A=1"
run c1d4 "MPET_P_CYCLES=1"
run c1d2 "MPET_P_CYCLES=1 MPET_P_DEGREE=2 MPET_P_DEGREE_COARSE=2"
run c2d2 "MPET_P_CYCLES=2 MPET_P_DEGREE=2 MPET_P_DEGREE_COARSE=2"
run c2d4c2 "MPET_P_CYCLES=2 MPET_P_DEGREE=4 MPET_P_DEGREE_COARSE=2"
run c1d4c2 "MPET_P_CYCLES=1 MPET_P_DEGREE=4 MPET_P_DEGREE_COARSE=2"
run c3d4c2 "MPET_P_CYCLES=3 MPET_P_DEGREE=4 MPET_P_DEGREE_COARSE=2"
run serial "MPET_PC_STREAMS=0 MPET_GRAPHS=0"
