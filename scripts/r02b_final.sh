#!/bin/bash
# round 2 (session 2): validation of the final state on one B200 -- full GPU suite, smoke, default bench line
mkdir -p gpurun_out
S=$SECONDS
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02b_final_tests.log 2>&1
echo "tests rc=$? $((SECONDS-S)) s"; tail -15 gpurun_out/r02b_final_tests.log | cut -c1-300
S=$SECONDS
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_final_smoke.log 2>&1
echo "smoke rc=$? $((SECONDS-S)) s"; tail -2 gpurun_out/r02b_final_smoke.log
S=$SECONDS
timeout 600 python bench.py > gpurun_out/r02b_final_bench.json 2> gpurun_out/r02b_final_bench.err
echo "bench rc=$? $((SECONDS-S)) s"; head -c 600 gpurun_out/r02b_final_bench.json; echo
