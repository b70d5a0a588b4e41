"""Which fp32 accumulation order do torch.matmul (4x4 @ 4x1 batched) and torch.cdist (mm route) use on this
device?  Compares ATen's result with candidate FMA / non-FMA chains emulated in float64.  Informational:
the kernels' choice (k-ordered FMA chain) is pinned by tests/test_gpu_stages.py."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcar_b200 import synthetic
f32 = np.float32
dev = "cuda" if torch.cuda.is_available() else "cpu"
def fma(a, b, c): return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)
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
cam = torch.matmul(mats, pts).squeeze(-1).cpu().numpy()[0]
Mn, pn = M.numpy(), p4.numpy()[0]
def chain(order, use_fma):
    out = np.zeros((6, 900, 4), f32)
    for c in range(6):
        for r in range(4):
            acc = None
            for k in order:
                a = np.full(900, Mn[c, r, k], f32); b = pn[:, k]
                if acc is None: acc = (a * b).astype(f32)
                elif use_fma: acc = fma(a, b, acc)
                else: acc = (acc + (a * b).astype(f32)).astype(f32)
            out[c, :, r] = acc
    return out
print("device", dev)
for order in ([0, 1, 2, 3], [3, 2, 1, 0]):
    for uf in (True, False):
        o = chain(order, uf); print("matmul4x4 order", order, "fma" if uf else "mul+add", "mismatch", int((o != cam).sum()), "of", cam.size)
x = (torch.rand(1, 900, 2, generator=g) * 100 - 50); y = (torch.rand(1, 1500, 2, generator=g) * 100 - 50)
d = torch.cdist(x.to(dev), y.to(dev), p=2.0)[0].cpu().numpy()
xn, yn = x[0].numpy(), y[0].numpy()
xnorm = ((xn[:, 0] * xn[:, 0]).astype(f32) + (xn[:, 1] * xn[:, 1]).astype(f32)).astype(f32)
ynorm = ((yn[:, 0] * yn[:, 0]).astype(f32) + (yn[:, 1] * yn[:, 1]).astype(f32)).astype(f32)
xnorm_f = fma(xn[:, 1], xn[:, 1], (xn[:, 0] * xn[:, 0]).astype(f32)); ynorm_f = fma(yn[:, 1], yn[:, 1], (yn[:, 0] * yn[:, 0]).astype(f32))
for nm, (xq, yq) in {"norm mul+add": (xnorm, ynorm), "norm fma": (xnorm_f, ynorm_f)}.items():
    A = np.stack([(-2 * xn[:, 0]).astype(f32), (-2 * xn[:, 1]).astype(f32), xq, np.ones(900, f32)], 1)
    Bm = np.stack([yn[:, 0], yn[:, 1], np.ones(1500, f32), yq], 1)
    for order in ([0, 1, 2, 3], [3, 2, 1, 0], [1, 0, 3, 2]):
        for uf in (True, False):
            acc = None
            for k in order:
                a = A[:, k][:, None].repeat(1500, 1); b = Bm[:, k][None, :].repeat(900, 0)
                if acc is None: acc = (a * b).astype(f32)
                elif uf: acc = fma(a, b, acc)
                else: acc = (acc + (a * b).astype(f32)).astype(f32)
            o = np.sqrt(np.maximum(acc, 0)).astype(f32)
            print("cdist", nm, order, "fma" if uf else "mul+add", "mismatch", int((o != d).sum()), "of", d.size, "maxdiff", float(np.abs(o - d).max()))
