"""Device time of the decoder self-attention launch (B=8, 900x900, 8 heads x 32) replayed back to back from a CUDA graph."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcar_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
f16 = "f16" in sys.argv            # fp16 operands + split-bf16 output: the bf16x3 mode's variant of the same kernel
q = torch.randn(B, 900, 768, device="cuda")
q = q.half() if f16 else q.bfloat16()
out = None
def run():
    ops.attention(q[:, :, :256], q[:, :, 256:512], q[:, :, 512:], 8, out_dtype="split" if f16 else None)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    run()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
reps = 20
with torch.cuda.graph(g):
    for _ in range(reps):
        run()
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    g.replay()
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) * 1e3 / (10 * reps)
flops = 4.0 * B * 8 * 900 * 900 * 32
print(f"self-attention B={B} {'fp16 / split out' if f16 else 'bf16'}: {t:.2f} us per launch ({flops / t * 1e-6:.1f} TFLOP/s, {B * 8 * 900 * 900 / t * 1e-3:.1f} G exp/s; MUFU floor {B * 8 * 900 * 900 / (148 * 16 * 1.965e3):.1f} us)")
