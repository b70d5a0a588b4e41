#!/bin/bash
# 8-GPU evidence: inference step (decode + NCCL gather in the timed region), training step (NCCL gradient all-reduce),
# VoVNet feature shapes (config 4), full model (config 3)
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "infer rc=$?"
timeout 600 $TR --master-port 29522 bench.py --gpus $N --config vovnet --steps 50 --warmup 5 > gpurun_out/bench_vovnet_${N}gpu.json 2> gpurun_out/bench_vovnet_${N}gpu.err; echo "vovnet rc=$?"
timeout 600 $TR --master-port 29523 bench.py --gpus $N --mode train --steps 20 --warmup 3 > gpurun_out/bench_train_${N}gpu.json 2> gpurun_out/bench_train_${N}gpu.err; echo "train rc=$?"
timeout 600 $TR --master-port 29524 bench.py --gpus $N --mode train --unfrozen --steps 10 --warmup 3 > gpurun_out/bench_train_unfrozen_${N}gpu.json 2> gpurun_out/bench_train_unfrozen_${N}gpu.err; echo "train-unfrozen rc=$?"
timeout 600 $TR --master-port 29525 bench.py --gpus $N --mode full --steps 10 --warmup 3 > gpurun_out/bench_full_${N}gpu.json 2> gpurun_out/bench_full_${N}gpu.err; echo "full rc=$?"
python - <<PY
import json
for f in ("bench_${N}gpu", "bench_vovnet_${N}gpu", "bench_train_${N}gpu", "bench_train_unfrozen_${N}gpu", "bench_full_${N}gpu"):
    try:
        d = [json.loads(l) for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1]
        extra = {k: d[k] for k in ("step", "train", "full") if k in d}
        e2e = {k: v for k, v in d.get("e2e", {}).items() if k != "note"}
        print(f, "value %.1f ms %.3f" % (d["value"], d["ms_per_step"]), extra, e2e)
    except Exception as exc:
        print(f, "FAILED", exc)
PY
