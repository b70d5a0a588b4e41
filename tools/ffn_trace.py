"""Per-CTA phase timeline of the fused feed-forward launch (tc_debug_trace hook; needs a TC_TRACE_BUILD library): `reps`
dependent blocks in one CUDA graph, median clock64 distances between the phase marks.  Usage: ffn_trace.py [reps]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcar_b200 import _lib, ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
M = 7200
dev = "cuda"
torch.manual_seed(0)
lib = _lib.load()
W1 = ops.mark_static(ops.cast_split(torch.randn(512, 256, device=dev) / 16)); W2 = ops.mark_static(ops.cast_split(torch.randn(256, 512, device=dev) / 22))
b1, b2 = torch.randn(512, device=dev) * 0.1, torch.randn(256, device=dev) * 0.1
ln = (torch.ones(256, device=dev), torch.zeros(256, device=dev))
x0 = torch.randn(M, 256, device=dev)
ctas = 2 * ((M + 127) // 128)
buf = torch.zeros(reps * ctas * 16, dtype=torch.int64, device=dev)
keep = []


def body():
    x32, x16 = x0, ops.cast_split(x0)
    for r in range(reps):
        x32, x16 = ops.ffn(x16, W1, b1, W2, b2, x32, ln)
        keep.append((x32, x16))


with torch.no_grad():
    body(); torch.cuda.synchronize()
    st = torch.cuda.Stream(); st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st): body()
    torch.cuda.current_stream().wait_stream(st)
    g = torch.cuda.CUDAGraph()
    lib.tc_debug_trace(buf.data_ptr(), reps * ctas)
    with torch.cuda.graph(g): body()
    lib.tc_debug_trace(None, 0)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(reps, ctas, 16).astype(np.int64)
g0, g9 = t[:, :, 0], t[:, :, 9]
base = g0[0].min()
for r in range(reps):
    print(f"  launch {r:2d}: first entry {(g0[r].min() - base) / 1e3:8.2f}  last entry {(g0[r].max() - base) / 1e3:8.2f}  "
          f"first exit {(g9[r].min() - base) / 1e3:8.2f}  last exit {(g9[r].max() - base) / 1e3:8.2f}  SMs {len(np.unique(t[r, :, 10]))}")
print(f"last-exit to last-exit: median {np.median(np.diff(g9[2:].max(1))) / 1e3:.2f} us")
names = [(1, "entry"), (3, "dependency wait returned"), (4, "GEMM 1: first stage landed"), (5, "GEMM 1: last stage landed"),
         (6, "hidden accumulator complete (conversion warps)"), (2, "GEMM 2: first stage ready (A tile converted + W2 landed)"),
         (7, "last A tile converted"), (11, "GEMM 2: last stage ready"), (12, "output accumulator complete"),
         (13, "cluster barrier 1 passed"), (14, "partials pushed + cluster barrier 2 passed"),
         (15, "rows final (LayerNorm statistics exchanged)"), (8, "exit (stores read, TMEM freed)")]
sel = t[2:]
print("median clock64 cycles from the dependency wait to each mark (per CTA, launches 2..):   [p10, median, p90]")
for slot, what in names:
    d = (sel[:, :, slot] - sel[:, :, 3]).ravel()
    print(f"  {what:58s} {np.percentile(d, 10):8.0f} {np.median(d):8.0f} {np.percentile(d, 90):8.0f}")
