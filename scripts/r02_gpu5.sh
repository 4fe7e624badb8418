#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_known_answer.py tests/test_gpu_api.py tests/test_gpu_solve.py tests/test_gpu_total_pressure.py tests/test_gpu_assembly.py tests/test_gpu_golden.py -q -s > gpurun_out/r02_gpu5_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_gpu5_tests.log
grep -n "passed\|failed\|FAILED\|Error\|iterations\|gpu .* oracle" gpurun_out/r02_gpu5_tests.log | cut -c1-300 | tail -25
