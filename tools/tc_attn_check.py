"""Quick GPU check of the tcgen05 attention path against fp64 (run under a timeout on the GPU box)."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcar_b200 import ops, synthetic
torch.manual_seed(0)
dev = "cuda"
def ref_attn(q, k, v, heads, blocked=None):
    B, Lq, E = q.shape; D = E // heads
    qh = q.double().view(B, Lq, heads, D).transpose(1, 2); kh = k.double().view(B, -1, heads, D).transpose(1, 2)
    vh = v.double().view(B, -1, heads, D).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(D)
    if blocked is not None: s = s.masked_fill(blocked.unsqueeze(1), float("-inf"))
    p = torch.nan_to_num(torch.softmax(s, -1), nan=0.0)
    return (p @ vh).transpose(1, 2).reshape(B, Lq, E).float()
for (B, Lq, Lk) in [(1, 128, 128), (1, 128, 256), (2, 900, 900), (1, 130, 77), (8, 900, 1500)]:
    q = torch.randn(B, Lq, 256, device=dev).bfloat16(); k = torch.randn(B, Lk, 256, device=dev).bfloat16(); v = torch.randn(B, Lk, 256, device=dev).bfloat16()
    out, _ = ops.attention(q, k, v, 8)
    torch.cuda.synchronize()
    want = ref_attn(q, k, v, 8)
    err = (out.float() - want).abs()
    print(f"dense B={B} Lq={Lq} Lk={Lk}: max err {err.max().item():.4e} mean {err.mean().item():.3e} |want| {want.abs().mean().item():.3f}", flush=True)
# masked
B, Q, R = 2, 900, 1500
g = torch.Generator().manual_seed(3)
radar_xy = (torch.rand((B, R, 2), generator=g) * 102.4 - 51.2); radar_xy[:, 1200:] = 500.0
centre = torch.rand((B, Q, 3), generator=g); code = torch.randn((B, Q, 10), generator=g) * 0.3; code[..., 3] += 1.0
radar_xy, centre, code = radar_xy.to(dev), centre.to(dev), code.to(dev)
geom = ops.radar_geometry(centre.view(B * Q, 3), code.view(B * Q, 10), synthetic.PC_RANGE, 1.0, 2.0, True)
blocked, row_any_ref = ops.radar_mask(geom, radar_xy, B, Q, R)
q = torch.randn(B, Q, 256, device=dev).bfloat16(); kv = torch.randn(B, R, 512, device=dev).bfloat16()
out, row_any = ops.attention(q, kv[:, :, :256], kv[:, :, 256:], 8, geom=geom, key_xy=radar_xy, want_row_any=True)
torch.cuda.synchronize()
want = ref_attn(q, kv[:, :, :256], kv[:, :, 256:], 8, blocked.bool())
err = (out.float() - want).abs()
print(f"masked: max err {err.max().item():.4e}; row_any equal: {torch.equal(row_any, row_any_ref)}; rows {int(row_any.sum())}")
# timing
q = torch.randn(8, 900, 768, device=dev).bfloat16()
for _ in range(3): ops.attention(q[:, :, :256], q[:, :, 256:512], q[:, :, 512:], 8)
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.attention(q[:, :, :256], q[:, :, 256:512], q[:, :, 512:], 8)
e1.record(); torch.cuda.synchronize()
print("self-attn B=8 900x900x8x32: %.1f us per call" % (e0.elapsed_time(e1) * 1e3 / 20))
