#!/bin/bash
# N-GPU evidence (gpurun --gpus N): inference step with decode + NCCL result gather in the timed region, VoVNet feature shapes,
# and the training step with the NCCL gradient all-reduce.  Usage: gpu_multi.sh N [infer|all]
mkdir -p gpurun_out
N=${1:-8}; WHAT=${2:-all}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "infer rc=$?"
if [ "$WHAT" = all ]; then
timeout 600 $TR --master-port 29522 bench.py --gpus $N --config vovnet --steps 50 --warmup 5 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench_vovnet_${N}gpu.json 2> gpurun_out/bench_vovnet_${N}gpu.err; echo "vovnet rc=$?"
timeout 600 $TR --master-port 29523 bench.py --gpus $N --mode train --steps 20 --warmup 3 > gpurun_out/bench_train_${N}gpu.json 2> gpurun_out/bench_train_${N}gpu.err; echo "train rc=$?"
fi
python - <<PY
import json
for f in ("bench_${N}gpu", "bench_vovnet_${N}gpu", "bench_train_${N}gpu"):
    try:
        d = [json.loads(l) for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1]
        e2e = {k: v for k, v in d.get("e2e", {}).items() if k != "note"}
        print(f, "value %.1f ms %.3f" % (d["value"], d["ms_per_step"]), d.get("step", d.get("train")), e2e)
    except Exception as exc:
        print(f, "FAILED", exc)
PY
