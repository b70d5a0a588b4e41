#!/bin/bash
# Turn the scratch outputs of tools/gpu_round.sh (gpurun_out/) into the tracked summaries under profiles/.
# Usage: bash tools/make_profiles.sh r02
R=${1:-r02}
mkdir -p profiles
python tools/launch_summary.py gpurun_out/launches.csv > profiles/${R}_launches_summary.txt
cp gpurun_out/launches.csv profiles/${R}_launches_step.csv
for k in sample attn linear attn_sparse ffn mlp; do
  [ -f gpurun_out/prof_$k.ncu-rep ] && python tools/ncu_metrics.py gpurun_out/prof_$k.ncu-rep pipe_xu.avg.pct_of_peak_sustained_active lts__t_bytes.sum l1tex__m_xbar2l1tex_read_bytes.sum lts__throughput.avg.pct > profiles/${R}_ncu_$k.txt
done
[ -f gpurun_out/prof_sample.ncu-rep ] && python tools/ncu_traffic.py gpurun_out/prof_sample.ncu-rep profiles/${R}_k1_traffic.json > /dev/null
[ -f gpurun_out/prof_sample.ncu-rep ] && python tools/ncu_source.py gpurun_out/prof_sample.ncu-rep 1.0 > profiles/${R}_ncu_sample_source_lines.txt
[ -f gpurun_out/prof_attn.ncu-rep ] && python tools/ncu_source.py gpurun_out/prof_attn.ncu-rep 1.0 > profiles/${R}_ncu_attn_source_lines.txt
grep -v "UserWarning\|_warn_once" gpurun_out/step_trace.txt > profiles/${R}_step_trace_cupti.txt
cp gpurun_out/step_timeline.log profiles/${R}_step_timeline.txt
for f in bench bench_ref bench_4gpu bench_bf16_onepass bench_vovnet bench_train_1gpu bench_train_unfrozen_1gpu bench_full_1gpu \
         bench_2gpu bench_8gpu bench_vovnet_8gpu bench_train_2gpu bench_train_8gpu bench_train_unfrozen_8gpu bench_full_8gpu; do
  [ -f gpurun_out/$f.json ] && grep "^{" gpurun_out/$f.json | tail -1 > profiles/${R}_$f.json
done
cat gpurun_out/k1_bench.log gpurun_out/attn_bench.log gpurun_out/linear_bench.log > profiles/${R}_kernel_microbench.txt
grep -E "passed|failed|\[" gpurun_out/pytest_gpu.log | grep -v "^=\+$" > profiles/${R}_pytest_gpu.txt
tail -2 gpurun_out/smoke.log > profiles/${R}_smoke.txt
python tools/sass_summary.py > profiles/${R}_sass_tc.txt
ls -la profiles | grep ${R}
