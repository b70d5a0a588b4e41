"""Isolated device time of the step's Linear shapes in the one-pass bf16 and the bf16x3 mode: `reps` back-to-back launches
(rotating over 4 operand sets, outputs kept distinct) in one CUDA graph between two events.  Usage: linear_bench.py [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcar_b200 import ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
dev = "cuda"
torch.manual_seed(0)
SHAPES = [  # (M, N, K, kind)
    (7200, 256, 256, "relu16"), (7200, 256, 256, "res+ln"), (7200, 512, 256, "relu16"), (7200, 256, 512, "res+ln"),
    (7200, 768, 256, "rb16"), (7200, 24, 256, "rb32"), (7200, 10, 256, "f32"), (12000, 1536, 256, "kv"),
]
def run(mode, M, N, K, kind):
    sets = []
    for i in range(4):
        A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * K ** -0.5
        A16, W16 = (ops.cast_split(A), ops.cast_split(W)) if mode == "x3" else (A.bfloat16(), W.bfloat16())
        b = torch.randn(N, device=dev) * 0.1
        kw = dict(out16="split" if mode == "x3" else "bf16")
        if kind == "relu16": kw.update(relu=True, want_f32=False, want_bf16=True)
        elif kind == "res+ln":
            kw.update(residual=torch.randn(M, N, device=dev), ln=(torch.ones(N, device=dev), torch.zeros(N, device=dev)), want_f32=True, want_bf16=True)
        elif kind == "rb16": kw.update(row_bias=torch.randn(900, N, device=dev), row_bias_period=900, want_f32=False, want_bf16=True,
                                        out16="f16" if mode == "x3" else "bf16"); b = None
        elif kind == "rb32": kw.update(row_bias=torch.randn(900, N, device=dev), row_bias_period=900); b = None
        elif kind == "f32": pass
        elif kind == "kv": kw.update(want_f32=mode == "x3", want_bf16=mode != "x3", out16="bf16")
        sets.append((A16, W16, b, kw))
    def body():
        for r in range(reps):
            A16, W16, b, kw = sets[r % 4]
            ops.linear(A16, W16, b, **kw)
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
    for M, N, K, kind in SHAPES:
        t1, t3 = run("bf16", M, N, K, kind), run("x3", M, N, K, kind)
        gf = 2 * M * N * K / 1e9
        print(f"M{M:6d} N{N:5d} K{K:4d} {kind:8s}  bf16 {t1:6.2f} us ({gf / t1 * 1e3:6.1f} TF/s)   bf16x3 {t3:6.2f} us ({gf / t3 * 1e3:6.1f} TF/s algorithmic)  x{t3 / t1:.2f}", flush=True)
