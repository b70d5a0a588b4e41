"""Fused three-layer head (tc_mlp) vs the three tc_linear launches it replaces: `reps` launches back to back in one CUDA graph
between two events (bf16x3 operands, M = 7200, C = 256, N3 = 10), plain ReLU heads and the LayerNorm (classification) variant.
Usage: mlp_bench.py [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcar_b200 import ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
M = 7200
dev = "cuda"
torch.manual_seed(0)
W1, W2 = (ops.mark_static(ops.cast_split(torch.randn(256, 256, device=dev) / 16)) for _ in range(2))
W3 = ops.mark_static(ops.cast_split(torch.randn(10, 256, device=dev) / 16))
b1, b2, b3 = torch.randn(256, device=dev) * 0.1, torch.randn(256, device=dev) * 0.1, torch.randn(10, device=dev) * 0.1
ln = (torch.ones(256, device=dev), torch.zeros(256, device=dev))
xs = [ops.cast_split(torch.randn(M, 256, device=dev)) for _ in range(4)]


def fused(x, use_ln):
    return ops.mlp(x, W1, b1, W2, b2, W3, b3, ln1=ln if use_ln else None, ln2=ln if use_ln else None)


def unfused(x, use_ln):
    kw = dict(ln=ln) if use_ln else {}
    _, a = ops.linear(x, W1, b1, relu=True, want_f32=False, want_bf16=True, out16="split", **kw)
    _, b = ops.linear(a, W2, b2, relu=True, want_f32=False, want_bf16=True, out16="split", **kw)
    return ops.linear(b, W3, b3)[0]


def run(fn, use_ln):
    keep = []
    def body():
        for r in range(reps):
            keep.append(fn(xs[r % 4], use_ln))
    body(); torch.cuda.synchronize()
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
    return best


with torch.no_grad():
    for use_ln in (False, True):
        print(f"M{M} three-layer head 256 -> 256 -> 256 -> 10{' with LayerNorms' if use_ln else ''}, bf16x3: fused {run(fused, use_ln):.2f} us, "
              f"three launches {run(unfused, use_ln):.2f} us", flush=True)
