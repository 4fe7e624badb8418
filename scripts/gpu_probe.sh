#!/bin/bash
mkdir -p gpurun_out
MPET_AMG_VERBOSE=1 timeout 600 python scripts/probe.py cfg5 72 1e-6 > gpurun_out/probe.log 2>&1
grep -E "amg\]|assemble_lhs|^spmv|step " gpurun_out/probe.log
