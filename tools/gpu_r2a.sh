#!/bin/bash
# round 2 GPU pass: all parity tests, quick bench of both tensor-core modes, in-graph timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -30
for prec in bf16x3 bf16; do
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --precision $prec > gpurun_out/bench_$prec.json 2> gpurun_out/bench_$prec.err; tail -3 gpurun_out/bench_$prec.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$prec.json')); r=d['roofline']; print('$prec samples/s', d['value'], 'ms', d['ms_per_step'], 'frac', r['frac'], 'k1 ms', r['avg_launch_ms'], 'iso', r.get('isolated',{}).get('avg_launch_ms'), 'e2e', d['e2e']['value'])"
done
timeout 300 python tools/step_timeline.py 20 bf16x3 > gpurun_out/step_timeline_x3.log 2>&1; head -40 gpurun_out/step_timeline_x3.log
