#!/bin/bash
# Quick GPU iteration: parity tests, short bench, launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 --cpu-samples 2 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 1 bf16 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
