#!/bin/bash
# Full ncu capture of selected kernels of one un-graphed step.  Usage: gpu_ncu.sh name:regex:skip:count ...
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read -r name regex skip count <<< "$spec"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k "regex:$regex" -s "$skip" -c "$count" -f -o "gpurun_out/prof_$name" python tools/profile_step.py 1 bf16x3 \
    > "gpurun_out/ncu_$name.log" 2>&1
  echo "ncu $name rc=$?"
done
ls -la gpurun_out/*.ncu-rep
