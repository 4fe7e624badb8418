#!/bin/bash
# round 2 (session 2), 1 GPU: is MINRES sensitive to how the SECOND network's block is inverted?  (the 8-GPU run lost it
# to a light V-cycle when its condition number crossed the polynomial threshold)
mkdir -p gpurun_out
run() {  # name, env
  env $2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02b_g3_$1.json 2> gpurun_out/r02b_g3_$1.err
  python - <<P
import json
for line in open('gpurun_out/r02b_g3_$1.json'):
    if line.startswith('{"metric"'):
        d=json.loads(line); r=d["roofline"]
        print("$1", round(d["ms_per_step"],1), d["config"]["krylov_iterations"], "ms/it", round(r["ms_per_iteration"],3), "pc", round(r["preconditioner"]["avg_application_ms"],3))
P
  grep "P1 field\|Error" gpurun_out/r02b_g3_$1.err | sort | uniq | head -6
}
run base "MPET_AMG_VERBOSE=1"
run f1_vcycle "MPET_AMG_VERBOSE=1 MPET_POLY_KAPPA_MAX=5.3"
run f1_vcycle_c2 "MPET_POLY_KAPPA_MAX=5.3 MPET_P_CYCLES=2"
run f1_vcycle_c2d4 "MPET_POLY_KAPPA_MAX=5.3 MPET_P_CYCLES=2 MPET_P_DEGREE=4 MPET_P_DEGREE_COARSE=4"
run all_vcycle "MPET_NO_POLY=1"
