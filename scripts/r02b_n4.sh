#!/bin/bash
# round 2 (session 2), 4 GPUs: does the iteration count depend on the partition?  Same global mesh (57^3 and 114^3) on
# 1 GPU / 4 bricks / 4 slabs
mkdir -p gpurun_out
show() { python - <<P
import json
for line in open('gpurun_out/r02b_n4_$1.json'):
    if line.startswith('{"metric"'):
        d=json.loads(line); r=d["roofline"]
        print("$1", round(d["ms_per_step"],1), d["config"]["krylov_iterations"], "ms/it", round(r["ms_per_iteration"],3), "pc", round(r["preconditioner"]["avg_application_ms"],3), "parity", (d["parity"] or {}).get("max_field_rel_err"), "res", d["true_residual"]["rel"])
P
  grep -i "error\|Traceback" gpurun_out/r02b_n4_$1.err | head -5
}
run4() {  # name, mesh-n, partition, env
  S=$SECONDS
  env $4 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 \
     bench.py --gpus 4 --steps 3 --warmup 3 --mesh-n $2 --partition $3 > gpurun_out/r02b_n4_$1.json 2> gpurun_out/r02b_n4_$1.err
  echo "$1 rc=$? $((SECONDS-S)) s"; show $1
}
S=$SECONDS
timeout 600 python bench.py --steps 3 --warmup 3 --mesh-n 57 --no-cpu-baseline --no-configs > gpurun_out/r02b_n4_single57.json 2> gpurun_out/r02b_n4_single57.err
echo "single57 rc=$? $((SECONDS-S)) s"; show single57
run4 brick36 36 brick A=1
run4 slab36 36 slab A=1
run4 brick72 72 brick A=1
run4 slab72 72 slab A=1
run4 brick72_strong 72 brick MPET_PC_LIGHT=0
