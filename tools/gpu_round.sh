#!/bin/bash
# Round evidence on the GPU box (via gpurun): parity tests, smoke, both bench arms, CUPTI timeline, ncu launch list and
# full captures of the sampling / attention / linear kernels.  Outputs land in gpurun_out/; summaries are made here by
# tools/make_profiles.sh and committed under profiles/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python tools/step_trace.py gpurun_out/step_trace.json > gpurun_out/step_trace.txt 2>&1; echo "trace rc=$?"
timeout 300 python tools/k1_bench.py > gpurun_out/k1_bench.log 2>&1; tail -1 gpurun_out/k1_bench.log
timeout 300 python tools/attn_bench.py 8 > gpurun_out/attn_bench.log 2>&1; tail -1 gpurun_out/attn_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 1 bf16 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
bash tools/gpu_ncu.sh sample:sample_kernel:2:1 attn:attention_tc:2:1 linear:linear_tc:12:9 attn_sparse:attention_sparse:1:1
ls -la gpurun_out
