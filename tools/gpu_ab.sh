#!/bin/bash
# A/B of environment switches on the default bench step.  Usage: gpu_ab.sh [reps] VAR=1 VAR2=1 ...   (each compared with the default)
mkdir -p gpurun_out
REPS=${1:-2}; shift
run() {
  env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('$*', 'samples/s %.0f fwd %.4f k1 %.2f us frac %.3f iso %.2f' % (d['value'], d['step']['forward_ms'], r['avg_launch_ms']*1e3, r['frac'], r['isolated']['avg_launch_ms']*1e3))" || tail -3 gpurun_out/ab.err
}
for i in $(seq $REPS); do
  run A=0
  for v in "$@"; do run $v; done
done
