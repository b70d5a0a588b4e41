"""Per-kernel device times INSIDE the captured step (warm L2, no host launch gaps): every library call is bracketed
by external CUDA events that become nodes of the CUDA graph.  Each pair adds ~1-2 us of its own, so read shares.
Usage: python tools/step_timeline.py [replays] [precision] [--seq]"""
import collections, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcar_b200 import ops, plugin, synthetic
replays = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20
precision = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "bf16x3"
B = 8
cfg = synthetic.head_config(900); cfg["precision"] = precision
head = plugin.build_head(cfg); head.load_state_dict(synthetic.make_state_dict(0, 900)); head = head.cuda().eval()
eng = head.engine()
dt = torch.float32 if precision == "fp32" else torch.bfloat16
feats = [f.to(dt).cuda().permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3) for f in synthetic.make_feats(0, B, "res101", smooth=True)]
prepared = eng.prepare_inputs(feats, synthetic.make_img_metas(B, seed=0))
with torch.no_grad():
    eng._forward_eager(prepared)
    torch.cuda.synchronize()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    ops.TIMELINE = []
    with torch.cuda.graph(g):
        out = eng._forward_eager(prepared)
    tl, ops.TIMELINE = ops.TIMELINE, None
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
acc = [0.0] * len(tl); total = 0.0
for it in range(replays + 2):
    flush.fill_(1)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(); g.replay(); s1.record(); torch.cuda.synchronize()
    if it >= 2:
        total += s0.elapsed_time(s1)
        for i, (_, a, b) in enumerate(tl):
            acc[i] += a.elapsed_time(b)
per = [x / replays * 1e3 for x in acc]
print(f"step (instrumented graph) {total / replays * 1e3:.1f} us, {len(tl)} launches, sum of brackets {sum(per):.1f} us")
agg = collections.OrderedDict()
for (label, _, _), t in zip(tl, per):
    a = agg.setdefault(label, [0, 0.0]); a[0] += 1; a[1] += t
for label, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:9.1f} us {n:3d} x {t / n:6.1f}  {label}")
if "--seq" in sys.argv:
    for (label, _, _), t in zip(tl, per):
        print(f"{t:7.1f}  {label}")
