#!/bin/bash
# 2-GPU pass: training parity (1 GPU), then the inference step with the NCCL result gather and the training step with the
# NCCL gradient all-reduce on 2 ranks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -q -x 2>&1 | tail -8
timeout 600 python bench.py --mode train --steps 20 --warmup 3 > gpurun_out/bench_train_1gpu.json 2> gpurun_out/bench_train_1gpu.err; echo "train1 rc=$?"; tail -2 gpurun_out/bench_train_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/bench_train_1gpu.json')); print('train 1gpu', d['value'], d['ms_per_step'], d['train'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "infer2 rc=$?"; tail -3 gpurun_out/bench_2gpu.err; python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu.json')); print('infer 2gpu', d['value'], d['ms_per_step'], d['step'], 'e2e', d['e2e']['value'], d['e2e']['h2d_gbs_per_rank'], d['e2e']['h2d_ceiling_gbs_per_rank'], d['e2e']['numa_node'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode train --steps 20 --warmup 3 > gpurun_out/bench_train_2gpu.json 2> gpurun_out/bench_train_2gpu.err; echo "train2 rc=$?"; tail -3 gpurun_out/bench_train_2gpu.err; python -c "
import json; d=json.load(open('gpurun_out/bench_train_2gpu.json')); print('train 2gpu', d['value'], d['ms_per_step'], d['train'])"
