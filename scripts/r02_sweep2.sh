#!/bin/bash
# round 2: displacement-cycle smoothing degree + light P1 cycles; ncu launch list of one step (1 GPU, cfg5)
mkdir -p gpurun_out
run() {  # name, env
  env $2 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_sw2_$1.log 2>&1
  python - <<P
import json
for line in open('gpurun_out/r02_sw2_$1.log'):
    if line.startswith('{"metric"'):
        d=json.loads(line); r=d["roofline"]
        print("$1", round(d["ms_per_step"],1), d["config"]["krylov_iterations"], "ms/it", round(r["ms_per_iteration"],3), "pc", round(r["preconditioner"]["avg_application_ms"],3), "launches/it", round(r["launches_per_iteration"]), "res", d["true_residual"]["rel"])
P
}
run light "MPET_P_CYCLES=1 MPET_P_DEGREE=2 MPET_P_DEGREE_COARSE=2"
run light_u1 "MPET_P_CYCLES=1 MPET_P_DEGREE=2 MPET_P_DEGREE_COARSE=2 MPET_U_DEGREE=1"
run light_u3 "MPET_P_CYCLES=1 MPET_P_DEGREE=2 MPET_P_DEGREE_COARSE=2 MPET_U_DEGREE=3"
run light_noprof "MPET_P_CYCLES=1 MPET_P_DEGREE=2 MPET_P_DEGREE_COARSE=2 MPET_BENCH_NOPROF=1"
# launch list: 1 timed step, eager launches (graphs off so that every kernel is a separate launch)
MPET_GRAPHS=0 MPET_P_CYCLES=1 MPET_P_DEGREE=2 MPET_P_DEGREE_COARSE=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/r02_launches_cfg5_light.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_sw2_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r02_launches_cfg5_light.csv 40 > gpurun_out/r02_launches_cfg5_light_summary.txt 2>&1
tail -45 gpurun_out/r02_launches_cfg5_light_summary.txt
