#!/bin/bash
# round 2 (session 2), 2 GPUs: multi-GPU parity tests, default bench line, serial breakdown of the preconditioner
mkdir -p gpurun_out
S=$SECONDS
timeout 600 python -m pytest tests/test_gpu_multi.py -q -s > gpurun_out/r02b_n2_tests.log 2>&1
echo "tests rc=$? $((SECONDS-S)) s"; tail -3 gpurun_out/r02b_n2_tests.log
run() {  # name, extra env
  S=$SECONDS
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
     bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02b_n2_$1.json 2> gpurun_out/r02b_n2_$1.err
  echo "$1 rc=$? $((SECONDS-S)) s"; python - <<P
import json
for line in open('gpurun_out/r02b_n2_$1.json'):
    if line.startswith('{"metric"'):
        d=json.loads(line); r=d["roofline"]
        print(d["ms_per_step"], d["config"]["krylov_iterations"], "ms/it", r["ms_per_iteration"], "spmv", r["avg_launch_ms"], "pc", r["preconditioner"]["avg_application_ms"], r["share_of_step"], "launches/it", r["launches_per_iteration"], "parity", d["parity"], "setup", d["config"]["amg_setup_s_excluded"], d["config"]["first_step_s"])
P
  grep -i "error\|Traceback\|mpet amg\|mpet dist" gpurun_out/r02b_n2_$1.err | head -20
}
run default "MPET_AMG_VERBOSE=1"
run serial "MPET_PC_STREAMS=0 MPET_GRAPHS=0"
