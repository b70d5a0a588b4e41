#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/linear_bench.py 24 2>&1 | tee gpurun_out/linear_bench.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:linear_tc" -s 20 -c 10 -f -o gpurun_out/prof_linear_x3 python tools/profile_step.py 1 bf16x3 > gpurun_out/ncu_linear_x3.log 2>&1; echo "ncu rc=$?"
