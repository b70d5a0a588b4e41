#!/bin/bash
# State check on the GPU box: parity tests, smoke, bench, chain microbench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python tools/chain_bench.py > gpurun_out/chain_bench.log 2>&1; echo "chain rc=$?"; cat gpurun_out/chain_bench.log
timeout 300 python tools/step_timeline.py > gpurun_out/step_timeline.log 2>&1; echo "timeline rc=$?"; tail -60 gpurun_out/step_timeline.log
