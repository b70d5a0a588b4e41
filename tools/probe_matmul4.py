"""Brute-force which evaluation tree torch.matmul([...,4,4] @ [...,4,1]) uses on this device (fp32).
Candidates: every permutation of the 4 products combined as (a) an FMA chain, (b) mul+add chain,
(c) two FMA pairs added, (d) two mul+add pairs added, (e) mixed."""
import itertools, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcar_b200 import synthetic
f32 = np.float32
dev = "cuda" if torch.cuda.is_available() else "cpu"
def fma(a, b, c): return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)
def mul(a, b): return (a * b).astype(f32)
def add(a, b): return (a + b).astype(f32)
calib = synthetic.load_calibration()
M = torch.tensor(calib[0], dtype=torch.float32)
g = torch.Generator().manual_seed(0)
ref = torch.rand(1, 900, 3, generator=g)
pc = synthetic.PC_RANGE
p = ref.clone()
for i in range(3): p[..., i] = p[..., i] * (pc[3 + i] - pc[i]) + pc[i]
p4 = torch.cat((p, torch.ones_like(p[..., :1])), -1)
pts = p4.view(1, 1, 900, 4).repeat(1, 6, 1, 1).unsqueeze(-1).to(dev)
mats = M.view(1, 6, 1, 4, 4).repeat(1, 1, 900, 1, 1).to(dev)
cam = torch.matmul(mats, pts).squeeze(-1).cpu().numpy()[0]           # [6,900,4]
Mn, pn = M.numpy(), p4.numpy()[0]
A = [np.broadcast_to(Mn[:, None, :, k], (6, 900, 4)).astype(f32) for k in range(4)]     # a_k[c,q,r]
B = [np.broadcast_to(pn[None, :, None, k], (6, 900, 4)).astype(f32) for k in range(4)]
best = []
for perm in itertools.permutations(range(4)):
    i, j, k, l = perm
    cands = {
        "fma-chain": fma(A[l], B[l], fma(A[k], B[k], fma(A[j], B[j], mul(A[i], B[i])))),
        "muladd-chain": add(add(add(mul(A[i], B[i]), mul(A[j], B[j])), mul(A[k], B[k])), mul(A[l], B[l])),
        "fma-pairs": add(fma(A[j], B[j], mul(A[i], B[i])), fma(A[l], B[l], mul(A[k], B[k]))),
        "muladd-pairs": add(add(mul(A[i], B[i]), mul(A[j], B[j])), add(mul(A[k], B[k]), mul(A[l], B[l]))),
        "fma-chain-from-pairsum": fma(A[l], B[l], add(fma(A[j], B[j], mul(A[i], B[i])), mul(A[k], B[k]))),
    }
    for name, o in cands.items():
        best.append((int((o != cam).sum()), name, perm))
best.sort()
print("device", dev, "total elements", cam.size)
for b in best[:12]:
    print(b)
