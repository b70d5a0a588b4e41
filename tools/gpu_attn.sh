#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], r['frac'], r['avg_launch_ms'], r.get('isolated'))"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_tc -s 2 -c 1 -f -o gpurun_out/prof_attn python tools/profile_step.py 1 bf16 > gpurun_out/ncu_attn.log 2>&1; echo "ncu rc=$?"
