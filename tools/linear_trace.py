"""Per-CTA phase timeline of the tensor-core Linear (tc_debug_trace hook): `reps` dependent launches of one shape back to back
(eager launches on one stream, programmatic dependent launch as in the step), then for the later launches the median
clock64 distance between the phase marks of a CTA and the globaltimer layout of the launches.
Usage: linear_trace.py [x3|bf16] [M N K kind] [reps]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcar_b200 import _lib, ops

mode = sys.argv[1] if len(sys.argv) > 1 else "x3"
M, N, K = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (7200, 256, 256)
kind = sys.argv[5] if len(sys.argv) > 5 else "relu16"
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 12
dev = "cuda"
torch.manual_seed(0)
lib = _lib.load()
sets = []
for i in range(4):
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * K ** -0.5
    A16, W16 = (ops.cast_split(A), ops.mark_static(ops.cast_split(W))) if mode == "x3" else (A.bfloat16(), ops.mark_static(W.bfloat16()))
    b = torch.randn(N, device=dev) * 0.1
    kw = dict(out16="split" if mode == "x3" else "bf16")
    if kind == "relu16": kw.update(relu=True, want_f32=False, want_bf16=True)
    elif kind == "res+ln":
        kw.update(residual=torch.randn(M, N, device=dev), ln=(torch.ones(N, device=dev), torch.zeros(N, device=dev)), want_f32=True, want_bf16=True)
    outs = {}
    if kw.get("want_bf16"):
        outs["out_bf16"] = ops.SplitBf16(torch.empty(M, 2 * N, device=dev, dtype=torch.bfloat16)) if mode == "x3" else torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    if kw.get("want_f32", True):
        outs["out_f32"] = torch.empty(M, N, device=dev)
    sets.append((A16, W16, b, kw, outs))
ctas = ((N + 63) // 64) * ((M + 127) // 128)
buf = torch.zeros(reps * ctas * 16, dtype=torch.int64, device=dev)


def body():
    for r in range(reps):
        A16, W16, b, kw, outs = sets[r % 4]
        ops.linear(A16, W16, b, **kw, **outs)


with torch.no_grad():
    body(); torch.cuda.synchronize()
    if os.environ.get("EAGER"):
        lib.tc_debug_trace(buf.data_ptr(), reps * ctas)
        body(); torch.cuda.synchronize()
        lib.tc_debug_trace(None, 0)
    else:                               # one CUDA graph, as in the step: the trace pointers are baked into the captured launches
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
g0 = t[:, :, 0]; g9 = t[:, :, 9]
base = g0[0].min()
print(f"{mode} M{M} N{N} K{K} {kind}: {ctas} CTAs per launch, {reps} launches; globaltimer (us from first entry)")
for r in range(reps):
    print(f"  launch {r:2d}: first entry {(g0[r].min() - base) / 1e3:8.2f}  last entry {(g0[r].max() - base) / 1e3:8.2f}  "
          f"first exit {(g9[r].min() - base) / 1e3:8.2f}  last exit {(g9[r].max() - base) / 1e3:8.2f}  "
          f"SMs {len(np.unique(t[r, :, 10]))}")
per = (g9[2:].max(1)[1:] - g9[2:].max(1)[:-1]) / 1e3
print(f"last-exit to last-exit: median {np.median(per):.2f} us")
names = {1: "entry", 2: "prologue done (barriers, TMEM alloc)", 3: "dependency wait returned", 4: "first stage landed (MMA thread)",
         5: "last stage landed", 11: "epilogue: side loads issued", 6: "epilogue: accumulator complete", 7: "epilogue done", 8: "exit (after TMEM free)"}
order = [1, 2, 3, 4, 5, 6, 7, 8]
sel = t[3:]          # steady state
print("median clock64 cycles between marks (per CTA, launches 3..):   [p10, median, p90]")
for a, b in zip(order[:-1], order[1:]):
    d = (sel[:, :, b] - sel[:, :, a]).ravel()
    print(f"  {names[a]:42s} -> {names[b]:36s} {np.percentile(d, 10):8.0f} {np.median(d):8.0f} {np.percentile(d, 90):8.0f}")
for a, b, what in ((6, 12, "accumulator complete -> first tcgen05.ld done"), (12, 13, "-> chunk 0 stored / LN: statistics ready"),
                   (13, 14, "-> chunk 1 stored / LN: statistics exchanged"), (14, 15, "-> TMA stores issued"), (15, 7, "-> TMA stores complete")):
    d = (sel[:, :, b] - sel[:, :, a]).ravel()
    print(f"    {what:50s} {np.percentile(d, 10):8.0f} {np.median(d):8.0f} {np.percentile(d, 90):8.0f}")
d = (sel[:, :, 11] - sel[:, :, 3]).ravel()
print(f"  wait returned -> epilogue side loads issued: median {np.median(d):.0f}")
d = (sel[:, :, 8] - sel[:, :, 1]).ravel()
print(f"  CTA lifetime: median {np.median(d):.0f} cycles;  entry -> wait returned: {np.median((sel[:, :, 3] - sel[:, :, 1]).ravel()):.0f}")
lastexit = g9.max(1); firstwait = None
# per launch: when (globaltimer) did its CTAs get past the dependency wait?  mark 3 is clock64; approximate with entry + (mark3 - mark1) cycles
w3 = g0 + ((t[:, :, 3] - t[:, :, 1]) / 1.965).astype(np.int64)
for r in range(3, min(reps, 8)):
    print(f"  launch {r}: previous last exit -> first CTA past its wait {(w3[r].min() - lastexit[r - 1]) / 1e3:6.2f} us, median CTA {(np.median(w3[r]) - lastexit[r - 1]) / 1e3:6.2f} us; "
          f"own first exit {(g9[r].min() - lastexit[r - 1]) / 1e3:6.2f}, last exit {(lastexit[r] - lastexit[r - 1]) / 1e3:6.2f}")
# how long after the previous launch's last exit does a CTA get past its wait (globaltimer has ~1 us granularity on some parts)
for r in range(3, min(reps, 7)):
    prev_last = g9[r - 1].max()
    e = (g0[r] - prev_last) / 1e3
    print(f"  launch {r}: CTA entries relative to previous launch's last exit: min {e.min():.2f} median {np.median(e):.2f} max {e.max():.2f} us; "
          f"entered before it: {(e < 0).mean() * 100:.0f} %")
# the stragglers of one steady-state launch: which phase makes them late?
r = min(reps - 1, 6)
order_exit = np.argsort(g9[r])
print(f"launch {r}: CTAs by exit time (us after previous launch's last exit); phases in cycles")
print("   cta  sm   entry   waitret    exit |  wait->first  first->last  last->acc  epilogue | CTAs on the same SM in this launch")
smids = t[r, :, 10]
for idx in list(order_exit[:4]) + list(order_exit[len(order_exit) // 2 - 2:len(order_exit) // 2 + 2]) + list(order_exit[-12:]):
    x = t[r, idx]
    same = int((smids == x[10]).sum())
    print(f"  {idx:4d} {x[10]:3d} {(x[0] - lastexit[r - 1]) / 1e3:7.2f} {(w3[r, idx] - lastexit[r - 1]) / 1e3:8.2f} {(x[9] - lastexit[r - 1]) / 1e3:7.2f} |"
          f" {x[4] - x[3]:10d} {x[5] - x[4]:11d} {x[6] - x[5]:10d} {x[7] - x[6]:9d} | {same}  last epilogue warp +{x[15] - x[7]} cyc, exit +{x[8] - x[15]} cyc, "
          f"entry->exit {x[8] - x[1]} cyc = {(x[9] - x[0]) / 1e3:.2f} us")
cnt = np.bincount(np.bincount(smids.astype(np.int64), minlength=148), minlength=4)
print("SMs by number of CTAs of this launch they ran:", {k: int(v) for k, v in enumerate(cnt) if v})
