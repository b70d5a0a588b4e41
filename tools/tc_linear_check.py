"""Quick GPU check of the tcgen05 linear path against fp64 (run under a timeout on the GPU box)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcar_b200 import ops, _lib
torch.manual_seed(0)
dev = "cuda"
def check(M, N, K, ln=False, extras=False):
    A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
    b = torch.randn(N, device=dev) * 0.1
    kw = {}
    ref = A.double() @ W.double().t() + b.double()
    if extras:
        res = torch.randn(M, N, device=dev); kw["residual"] = res; ref = ref + res.double()
    if ln:
        g = 1 + 0.1 * torch.randn(N, device=dev); be = 0.1 * torch.randn(N, device=dev); kw["ln"] = (g, be)
        ref = torch.nn.functional.layer_norm(ref, (N,), g.double(), be.double(), 1e-5)
    kw["relu"] = True; ref = ref.relu()
    o32, o16 = ops.linear(A, W, b, want_f32=True, want_bf16=True, **kw)
    torch.cuda.synchronize()
    err = (o32.double() - ref).abs().max().item(); err16 = (o16.double() - ref).abs().max().item()
    print(f"M={M} N={N} K={K} ln={ln} extras={extras}: max err fp32-out {err:.3e} bf16-out {err16:.3e}", flush=True)
    return err
n0 = _lib.launch_count()
worst = 0
for shp in [(128, 64, 64), (128, 256, 64), (128, 256, 256), (900, 256, 256), (7200, 256, 256), (7200, 768, 256), (7200, 256, 512), (77, 512, 256), (130, 128, 64), (12000, 1536, 256)]:
    worst = max(worst, check(*shp))
worst = max(worst, check(7200, 256, 256, ln=True, extras=True), check(900, 256, 512, ln=True, extras=True))
print("worst", worst, "launches", _lib.launch_count() - n0)
# timing
A = torch.randn(7200, 256, device=dev).bfloat16(); W = torch.randn(256, 256, device=dev).bfloat16(); b = torch.zeros(256, device=dev)
o16 = torch.empty(7200, 256, device=dev, dtype=torch.bfloat16)
for _ in range(5): ops.linear(A, W, b, out_bf16=o16, want_f32=False)
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): ops.linear(A, W, b, out_bf16=o16, want_f32=False)
e1.record(); torch.cuda.synchronize()
print("7200x256x256 tc linear: %.2f us per call (incl. launch gaps)" % (e0.elapsed_time(e1) * 1e3 / 50))
assert worst < 1e-3
