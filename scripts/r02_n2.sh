#!/bin/bash
# round 2: 2 GPUs -- multi-GPU parity tests, then bench with the peer-memory channel vs NCCL halos
mkdir -p gpurun_out
export MPET_AMG_VERBOSE=1
timeout 600 python -m pytest tests/test_gpu_multi.py -q -s > gpurun_out/r02_n2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_n2_tests.log
tail -8 gpurun_out/r02_n2_tests.log
unset MPET_AMG_VERBOSE
run() {  # name, extra env, n
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 3 --warmup 3 --mesh-n $3 > gpurun_out/r02_n2_$1.log 2>&1
  echo "$1 rc=$?"; grep -o '"ms_per_step": [0-9.]*\|"krylov_iterations": \[[^]]*\]\|"max_field_rel_err": [^,]*\|"ms_per_iteration": [0-9.]*\|"parallelism": "[^"]*"\|"halo_and_collectives_main_stream": [0-9.e-]*' gpurun_out/r02_n2_$1.log | head -12
  tail -c 600 gpurun_out/r02_n2_$1.log | grep -i "error\|Traceback" | head
}
run peer36 "A=1" 36
run nccl36 "MPET_COMM=nccl" 36
run peer72 "A=1" 72
