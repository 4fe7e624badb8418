#!/bin/bash
# round 2 (session 2), call 1 (1 GPU): full GPU suite, the default bench line, launch list + ncu --set full of the shipped kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02b_smi.txt 2>&1; nproc >> gpurun_out/r02b_smi.txt
S=$SECONDS
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r02b_tests.log 2>&1
echo "tests rc=$? $((SECONDS-S)) s" | tee -a gpurun_out/r02b_tests.log
grep -n "passed\|failed\|FAILED\|Error" gpurun_out/r02b_tests.log | cut -c1-300 | tail -12
S=$SECONDS
timeout 1200 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
echo "bench rc=$? $((SECONDS-S)) s"
S=$SECONDS
timeout 600 python bench.py --impl reference > gpurun_out/r02b_bench_ref.json 2> gpurun_out/r02b_bench_ref.err
echo "ref bench rc=$? $((SECONDS-S)) s"
S=$SECONDS
MPET_GRAPHS=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
  --log-file gpurun_out/r02_launches_cfg5.csv python scripts/profile_step.py > gpurun_out/r02b_ncu1.log 2>&1
echo "launch list rc=$? $((SECONDS-S)) s"
python scripts/launch_summary.py gpurun_out/r02_launches_cfg5.csv 40 > gpurun_out/r02_launches_cfg5_summary.txt 2>&1
S=$SECONDS
MPET_GRAPHS=0 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"k_block_rows_pipe|k_spmm_pipe|k_spmm_x16" -c 40 -o gpurun_out/r02_spmv_spmm_full python scripts/profile_step.py > gpurun_out/r02b_ncu2.log 2>&1
echo "full capture rc=$? $((SECONDS-S)) s"
ls -la gpurun_out | tail -12
head -c 1500 gpurun_out/r02b_bench.json; echo
tail -30 gpurun_out/r02_launches_cfg5_summary.txt
