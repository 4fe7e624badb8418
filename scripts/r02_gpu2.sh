#!/bin/bash
# round 2, call 2: full GPU test suite (no -x) + launch list of one bench step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_gpu2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_gpu2_tests.log
grep -n "passed\|failed\|FAILED\|rates\|u_L2\|worst" gpurun_out/r02_gpu2_tests.log | tail -30
