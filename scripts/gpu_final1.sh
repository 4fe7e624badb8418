#!/bin/bash
# One-GPU round-end visit: parity tests, smoke, bench (both formulations), reference arm, launch list + ncu captures.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -1 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 900 python bench.py --formulation total-pressure --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tp.json 2> gpurun_out/bench_tp.err; cat gpurun_out/bench_tp.json; tail -2 gpurun_out/bench_tp.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --rtol 1e-2 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_block_rows_pipeILi3ELi3ELb1E -s 6 -c 1 -f -o gpurun_out/full_block_rows_pipe \
    python bench.py --steps 1 --warmup 3 --rtol 1e-2 --no-cpu-baseline > gpurun_out/ncu_full_blk.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_spmm_pipeILi4ELi2ELi1E -s 8 -c 1 -f -o gpurun_out/full_spmm_pipe_p2 \
    python bench.py --steps 1 --warmup 3 --rtol 1e-2 --no-cpu-baseline > gpurun_out/ncu_full_spmm.log 2>&1
ls -la gpurun_out/*.ncu-rep
