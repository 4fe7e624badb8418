#!/bin/bash
# round 2 (session 2), 8 GPUs: weak-scaling line after the polynomial-threshold fix; cfg4 (A=4, 71-72^3 cubes, ~10.3 M dofs)
# strong-scaled onto 4 and 8 GPUs
mkdir -p gpurun_out
show() { python - <<P
import json
for line in open('gpurun_out/r02b_n8b_$1.json'):
    if line.startswith('{"metric"'):
        d=json.loads(line); r=d["roofline"]
        print("$1", round(d["ms_per_step"],1), d["config"]["krylov_iterations"], "dofs", d["config"]["total_dofs"], "value", round(d["value"]), "ms/it", round(r["ms_per_iteration"],3), "spmv", round(r["avg_launch_ms"],3), "pc", round(r["preconditioner"]["avg_application_ms"],3), "parity", (d["parity"] or {}).get("max_field_rel_err"), "res", d["true_residual"]["rel"], "setup", d["config"]["amg_setup_s_excluded"])
P
  grep -i "error\|Traceback" gpurun_out/r02b_n8b_$1.err | head -5
}
runN() {  # name, nranks, extra bench args
  S=$SECONDS
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29519 \
     bench.py --gpus $2 --steps 3 --warmup 3 $3 > gpurun_out/r02b_n8b_$1.json 2> gpurun_out/r02b_n8b_$1.err
  echo "$1 rc=$? $((SECONDS-S)) s"; show $1
}
MPET_AMG_VERBOSE=1 runN weak8 8 ""
grep "P1 field" gpurun_out/r02b_n8b_weak8.err | sort | uniq -c | sort -rn | head -4
runN cfg4_strong8 8 "--config cfg4 --mesh-n 36"
runN cfg4_strong4 4 "--config cfg4 --mesh-n 45"
