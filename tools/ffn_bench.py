"""Fused feed-forward launch (tc_ffn) vs the two tc_linear launches it replaces: `reps` dependent blocks back to back in one
CUDA graph between two events (bf16x3 operands, M = 7200, C = 256, H = 512).  Usage: ffn_bench.py [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcar_b200 import ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
M = int(sys.argv[2]) if len(sys.argv) > 2 else 7200
dev = "cuda"
torch.manual_seed(0)
sets = []
for i in range(4):
    W1 = ops.mark_static(ops.cast_split(torch.randn(512, 256, device=dev) / 16)); W2 = ops.mark_static(ops.cast_split(torch.randn(256, 512, device=dev) / 22))
    b1, b2 = torch.randn(512, device=dev) * 0.1, torch.randn(256, device=dev) * 0.1
    ln = (torch.ones(256, device=dev), torch.zeros(256, device=dev))
    sets.append((W1, b1, W2, b2, ln))
x0 = torch.randn(M, 256, device=dev)


def fused(x32, x16, st):
    W1, b1, W2, b2, ln = st
    return ops.ffn(x16, W1, b1, W2, b2, x32, ln)


def unfused(x32, x16, st):
    W1, b1, W2, b2, ln = st
    _, h = ops.linear(x16, W1, b1, relu=True, want_f32=False, want_bf16=True, out16="split")
    return ops.linear(h, W2, b2, residual=x32, ln=ln, want_bf16=True, out16="split")


def run(fn):
    keep = []
    def body():
        x32, x16 = x0, ops.cast_split(x0)
        for r in range(reps):
            x32, x16 = fn(x32, x16, sets[r % 4])
            keep.append((x32, x16))
        return x32
    out = body(); torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): body()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): body()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    return best, out


with torch.no_grad():
    tf, of = run(fused)
    tu, ou = run(unfused)
    print(f"M{M} FFN block (256 -> 512 -> 256 + residual + LN), bf16x3: fused {tf:.2f} us, two launches {tu:.2f} us; "
          f"max |diff| after {reps} chained blocks {float((of - ou).abs().max()):.2e}")
