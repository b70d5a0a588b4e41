#!/bin/bash
# K1 tuning on the GPU box: the sampling kernel alone (all variants), then parity tests and a short bench.
mkdir -p gpurun_out
for v in ${K1_VARIANTS:-0 1 2 3 4}; do TC_SAMPLE_VARIANT=$v timeout 200 python tools/k1_bench.py 2>&1 | tail -1; done | tee gpurun_out/k1_bench.log
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
