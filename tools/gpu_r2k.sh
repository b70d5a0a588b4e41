#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -12
for prec in bf16x3 bf16; do
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-gpu-eager-baseline --precision $prec > gpurun_out/bench_$prec.json 2> gpurun_out/bench_$prec.err; tail -3 gpurun_out/bench_$prec.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$prec.json')); r=d['roofline']; print('$prec samples/s', d['value'], 'ms', d['ms_per_step'], 'fwd', d['step']['forward_ms'], 'frac', r['frac'], 'k1 ms', r['avg_launch_ms'], 'iso', r.get('isolated',{}).get('avg_launch_ms'), 'e2e', d['e2e']['value'], 'launches', d['gpu_launches_per_step'])"
done
timeout 300 python tools/step_trace.py gpurun_out/step_trace_x3.json bf16x3 > gpurun_out/step_trace_x3.txt 2>&1; head -3 gpurun_out/step_trace_x3.txt | tail -1
