#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_known_answer.py tests/test_gpu_solve.py tests/test_gpu_api.py tests/test_gpu_total_pressure.py tests/test_gpu_configs.py -q -s -x > gpurun_out/r02_gpu4_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_gpu4_tests.log
grep -n "passed\|failed\|FAILED\|Error\|iterations\|gpu .* oracle" gpurun_out/r02_gpu4_tests.log | cut -c1-300 | tail -25
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-n 20 --no-configs > gpurun_out/r02_gpu4_bench.log 2>&1
echo "bench rc=$?"
python - <<P
import json
for line in open('gpurun_out/r02_gpu4_bench.log'):
    if line.startswith('{"metric"'):
        d=json.loads(line); r=d["roofline"]
        print(round(d["ms_per_step"],1), d["config"]["krylov_iterations"], "e2e", round(d["e2e"]["ms_per_step"],1), "ms/it", round(r["ms_per_iteration"],3), "spmv", round(r["avg_launch_ms"],3), r["frac"], "pc", round(r["preconditioner"]["avg_application_ms"],3), r["preconditioner"]["frac"], "res", d["true_residual"]["rel"], "parity", d["parity"], "matched", {k:v for k,v in d["matched"].items() if k!="cpu_trend"})
P
tail -3 gpurun_out/r02_gpu4_bench.log | cut -c1-300
