#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/step_trace.py gpurun_out/step_trace_x3.json bf16x3 > gpurun_out/step_trace_x3.txt 2>&1; echo "trace rc=$?"; head -3 gpurun_out/step_trace_x3.txt
timeout 600 python bench.py --mode train --steps 20 --warmup 3 > gpurun_out/bench_train_1gpu.json 2> gpurun_out/bench_train_1gpu.err; echo "train rc=$?"; tail -2 gpurun_out/bench_train_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/bench_train_1gpu.json')); print('train 1gpu', d['value'], d['ms_per_step'], d['gpu_launches_per_step'], d['train']['exposed_allreduce_ms'])"
timeout 600 python bench.py --mode train --unfrozen --steps 10 --warmup 3 > gpurun_out/bench_train_unfrozen_1gpu.json 2> gpurun_out/bench_train_unfrozen_1gpu.err; echo "train-unfrozen rc=$?"; tail -2 gpurun_out/bench_train_unfrozen_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/bench_train_unfrozen_1gpu.json')); print('train unfrozen 1gpu', d['value'], d['ms_per_step'], d['gpu_launches_per_step'])"
