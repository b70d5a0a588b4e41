#!/bin/bash
# parity tests + short bench (no CPU baseline) + summary line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); r=d['roofline']; print('samples/s', d['value'], 'ms', d['ms_per_step'], 'frac', r['frac'], 'k1 ms', r['avg_launch_ms'], 'iso', r.get('isolated',{}).get('avg_launch_ms'))"
if [ -n "$TIMELINE" ]; then timeout 300 python tools/step_timeline.py > gpurun_out/step_timeline.log 2>&1; head -30 gpurun_out/step_timeline.log; fi
