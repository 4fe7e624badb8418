#!/bin/bash
# round 2: 2 GPUs, n=72 -- where does the preconditioner time go (peer channel on)
mkdir -p gpurun_out
run() {  # name, extra env
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
     bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_n2b_$1.log 2>&1
  echo "$1 rc=$?"; python - <<P
import json
for line in open('gpurun_out/r02_n2b_$1.log'):
    if line.startswith('{"metric"'):
        d=json.loads(line); r=d["roofline"]
        print(d["ms_per_step"], d["config"]["krylov_iterations"], "ms/it", r["ms_per_iteration"], "spmv", r["avg_launch_ms"], "pc", r["preconditioner"]["avg_application_ms"], r["share_of_step"], "launches/it", r["launches_per_iteration"], "parity", d["parity"].get("max_field_rel_err"))
P
  grep -i "error\|Traceback" gpurun_out/r02_n2b_$1.log | head -3
}
run default "A=1"
run serial "MPET_PC_STREAMS=0 MPET_GRAPHS=0"
run nographs "MPET_GRAPHS=0"
run pcyc2 "MPET_P_CYCLES=2"
