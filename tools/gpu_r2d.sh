#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 50 --warmup 5 --cpu-samples 3 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_r2d.err; cat gpurun_out/bench_r2d.json
timeout 600 python bench.py --mode train --steps 20 --warmup 3 > gpurun_out/bench_train_r2d.json 2> gpurun_out/bench_train_r2d.err; echo "train rc=$?"; tail -5 gpurun_out/bench_train_r2d.err; cat gpurun_out/bench_train_r2d.json
