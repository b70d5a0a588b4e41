import os, sys, torch
sys.path.insert(0, "/root/repo")
from transcar_b200 import ops
q = torch.randn(8, 900, 768, device="cuda").bfloat16()
for _ in range(2): ops.attention(q[:, :, :256], q[:, :, 256:512], q[:, :, 512:], 8)
torch.cuda.synchronize()
