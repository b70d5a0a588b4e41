#!/bin/bash
# Round evidence on the GPU box (via gpurun): parity tests, smoke, both bench arms, CUPTI timeline, in-graph timeline, kernel
# micro-benchmarks, ncu launch list and full captures of the sampling / attention / linear / sparse kernels (bf16x3 step).
# Outputs land in gpurun_out/; summaries are made here by tools/make_profiles.sh and committed under profiles/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --steps 200 --warmup 10 --precision bf16 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench_bf16_onepass.json 2> gpurun_out/bench_bf16_onepass.err; echo "bench one-pass rc=$?"
timeout 600 python bench.py --steps 100 --warmup 10 --config vovnet --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench_vovnet.json 2> gpurun_out/bench_vovnet.err; echo "bench vovnet rc=$?"
timeout 600 python bench.py --mode train --steps 20 --warmup 3 > gpurun_out/bench_train_1gpu.json 2> gpurun_out/bench_train_1gpu.err; echo "train rc=$?"
timeout 600 python bench.py --mode train --unfrozen --steps 10 --warmup 3 > gpurun_out/bench_train_unfrozen_1gpu.json 2> gpurun_out/bench_train_unfrozen_1gpu.err; echo "train unfrozen rc=$?"
timeout 600 python bench.py --mode full --steps 10 --warmup 3 > gpurun_out/bench_full_1gpu.json 2> gpurun_out/bench_full_1gpu.err; echo "full rc=$?"
timeout 300 python tools/step_trace.py gpurun_out/step_trace.json bf16x3 > gpurun_out/step_trace.txt 2>&1; echo "trace rc=$?"
timeout 300 python tools/step_timeline.py 20 bf16x3 > gpurun_out/step_timeline.log 2>&1; echo "timeline rc=$?"
timeout 300 python tools/k1_bench.py > gpurun_out/k1_bench.log 2>&1; tail -1 gpurun_out/k1_bench.log
(timeout 300 python tools/attn_bench.py 8; timeout 300 python tools/attn_bench.py 8 f16) > gpurun_out/attn_bench.log 2>&1; tail -2 gpurun_out/attn_bench.log
(timeout 300 python tools/linear_bench.py 24; timeout 200 python tools/ffn_bench.py 24; timeout 200 python tools/mlp_bench.py 24; timeout 300 python tools/sparse_attn_bench.py 20 2>&1 | grep "radar layer"; timeout 200 python tools/replay_cost.py 2>&1 | tail -2) > gpurun_out/linear_bench.log 2>&1; tail -5 gpurun_out/linear_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 1 bf16x3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
bash tools/gpu_ncu.sh sample:sample_kernel:2:1 attn:attention_tc:2:1 linear:linear_tc:12:9 attn_sparse:attention_sparse:1:1 ffn:ffn_tc:2:1 mlp:mlp_tc:2:2
ls -la gpurun_out | head -60
