#!/bin/bash
# Turn the scratch outputs of tools/gpu_round.sh (gpurun_out/) into the tracked summaries under profiles/.
# Usage: bash tools/make_profiles.sh r01
R=${1:-r01}
mkdir -p profiles
python tools/launch_summary.py gpurun_out/launches.csv > profiles/${R}_launches_summary.txt
cp gpurun_out/launches.csv profiles/${R}_launches_step.csv
for k in sample attn linear attn_sparse; do
  [ -f gpurun_out/prof_$k.ncu-rep ] && python tools/ncu_metrics.py gpurun_out/prof_$k.ncu-rep pipe_xu.avg.pct_of_peak_sustained_active lts__t_bytes.sum > profiles/${R}_ncu_$k.txt
done
[ -f gpurun_out/prof_sample.ncu-rep ] && python tools/ncu_traffic.py gpurun_out/prof_sample.ncu-rep profiles/${R}_k1_traffic.json > /dev/null
[ -f gpurun_out/prof_sample.ncu-rep ] && python tools/ncu_source.py gpurun_out/prof_sample.ncu-rep 1.0 > profiles/${R}_ncu_sample_source_lines.txt
[ -f gpurun_out/prof_attn.ncu-rep ] && python tools/ncu_source.py gpurun_out/prof_attn.ncu-rep 1.0 > profiles/${R}_ncu_attn_source_lines.txt
grep -v "UserWarning\|_warn_once" gpurun_out/step_trace.txt > profiles/${R}_step_trace_cupti.txt
cp gpurun_out/bench.json profiles/${R}_bench_1gpu.json
cp gpurun_out/bench_ref.json profiles/${R}_bench_reference_arm.json
cat gpurun_out/k1_bench.log gpurun_out/attn_bench.log > profiles/${R}_kernel_microbench.txt
ls -la profiles
