#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -12
timeout 600 python bench.py --mode full --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "full rc=$?"; tail -3 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
timeout 600 python bench.py --config vovnet --steps 50 --warmup 5 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench_vovnet.json 2> gpurun_out/bench_vovnet.err; echo "vovnet rc=$?"; tail -3 gpurun_out/bench_vovnet.err; python -c "
import json; d=json.load(open('gpurun_out/bench_vovnet.json')); r=d['roofline']; print('vovnet samples/s', d['value'], 'ms', d['ms_per_step'], 'frac', r['frac'], 'k1 ms', r['avg_launch_ms'], 'iso', r.get('isolated',{}).get('avg_launch_ms'))"
timeout 300 python tools/linear_bench.py 24 2>&1 | tee gpurun_out/linear_bench.log
