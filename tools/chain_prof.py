"""One fused decoder-tail chain launch (plus warm-up) for ncu: python tools/chain_prof.py [M]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from transcar_b200 import ops  # noqa: E402
import test_gpu_chain as T  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 7200
p = T._decoder_tail_case(M)
dev = p["s"].device
x2_32 = torch.empty((M, 256), device=dev)
x2_16 = torch.empty((M, 256), device=dev, dtype=torch.bfloat16)
code = torch.empty((M, 10), device=dev)
new_ref = torch.empty((M, 3), device=dev)
stages = T.decoder_tail_stages(ops, p, x2_32, x2_16, code, new_ref)
for _ in range(3):
    ops.linear_chain(p["s"], stages)
torch.cuda.synchronize()
