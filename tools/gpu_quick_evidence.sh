#!/bin/bash
# Minimal evidence refresh on the GPU box: parity tests, smoke, the default bench line (bf16x3, batch 8) and the fused-kernel
# micro-benchmarks.  Outputs land in gpurun_out/ (tools/make_profiles.sh copies the summaries to profiles/).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
(timeout 200 python tools/ffn_bench.py 24; timeout 300 python tools/linear_bench.py 24) > gpurun_out/linear_bench.log 2>&1; tail -4 gpurun_out/linear_bench.log
