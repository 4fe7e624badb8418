#!/bin/bash
# round 2 (session 2), call 2 (1 GPU): device-resident GMRES tests, default bench line, launch list, ncu --set full of the shipped kernels
mkdir -p gpurun_out
S=$SECONDS
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_solve.py tests/test_gpu_known_answer.py -m gpu -q -x > gpurun_out/r02b2_tests.log 2>&1
echo "tests rc=$? $((SECONDS-S)) s"; tail -3 gpurun_out/r02b2_tests.log
S=$SECONDS
timeout 1200 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
echo "bench rc=$? $((SECONDS-S)) s"
S=$SECONDS
MPET_GRAPHS=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
  --log-file /tmp/r02_launches_cfg5.csv python scripts/profile_step.py > gpurun_out/r02b_ncu1.log 2>&1
echo "launch list rc=$? $((SECONDS-S)) s"
python scripts/launch_summary.py /tmp/r02_launches_cfg5.csv 60 > gpurun_out/r02_launches_cfg5_summary.txt 2>&1
gzip -c /tmp/r02_launches_cfg5.csv > gpurun_out/r02_launches_cfg5.csv.gz
S=$SECONDS
MPET_GRAPHS=0 MPET_PC_STREAMS=0 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"k_block_rows_pipe|k_spmm_pipe" -c 12 -o /tmp/r02_spmv_spmm_full python scripts/profile_step.py > gpurun_out/r02b_ncu2.log 2>&1
echo "full capture rc=$? $((SECONDS-S)) s"
ncu -i /tmp/r02_spmv_spmm_full.ncu-rep --page raw --csv > gpurun_out/r02_spmv_spmm_full_raw.csv 2>/dev/null
ls -la /tmp/r02_spmv_spmm_full.ncu-rep
sz=$(stat -c %s /tmp/r02_spmv_spmm_full.ncu-rep)
if [ "$sz" -lt 30000000 ]; then cp /tmp/r02_spmv_spmm_full.ncu-rep gpurun_out/; fi
du -sh gpurun_out
head -32 gpurun_out/r02_launches_cfg5_summary.txt
