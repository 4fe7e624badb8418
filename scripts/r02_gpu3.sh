#!/bin/bash
# round 2, call 3 (1 GPU): full GPU suite + graphs on/off at N=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_gpu3_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_gpu3_tests.log
grep -n "passed\|failed\|FAILED\|standard\|total_pressure 0\|total_pressure 1\|worst" gpurun_out/r02_gpu3_tests.log | cut -c1-700 | tail -30
for g in 1 0; do
  MPET_GRAPHS=$g timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_gpu3_bench_g$g.log 2>&1
  echo "graphs=$g rc=$?"; grep -o '"ms_per_step": [0-9.]*\|"krylov_iterations": \[[^]]*\]\|"ms_per_iteration": [0-9.]*\|"avg_application_ms": [0-9.]*\|"avg_launch_ms": [0-9.]*' gpurun_out/r02_gpu3_bench_g$g.log | head -8
done
for c in cfg1 cfg2 cfg3; do for g in 1 0; do
  MPET_GRAPHS=$g timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_gpu3_${c}_g$g.log 2>&1
  echo "$c graphs=$g rc=$?"; grep -o '"ms_per_step": [0-9.]*\|"krylov_iterations": \[[^]]*\]\|"ms_per_iteration": [0-9.]*' gpurun_out/r02_gpu3_${c}_g$g.log | head -3
done; done
