"""Device time of the K1 sampling kernel alone on the bench workload (res101, batch 8): the six per-layer launches of one
step (their real reference points / logits, captured from an eager forward) replayed back to back from a CUDA graph.
Usage: [TC_SAMPLE_VARIANT=n] python tools/k1_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcar_b200 import ops, plugin, synthetic  # noqa: E402

mode = "cta-dynamic"
B = 8
cfg = synthetic.head_config(900)
cfg["precision"] = "bf16"
head = plugin.build_head(cfg)
head.load_state_dict(synthetic.make_state_dict(0, 900))
head = head.cuda().eval()
eng = head.engine()
eng.use_graph = False
feats = [f.to(torch.bfloat16).cuda().permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
         for f in synthetic.make_feats(0, B, "res101", smooth=False)]
prepared = eng.prepare_inputs(feats, synthetic.make_img_metas(B, seed=0))
calls = []
orig = ops.sample_fwd


def spy(*a, **k):
    calls.append((a, dict(k)))
    return orig(*a, **k)


ops.sample_fwd = spy
eng.forward_prepared(prepared)
torch.cuda.synchronize()
ops.sample_fwd = orig
outs = [torch.empty((B, 900, 256), device="cuda", dtype=torch.bfloat16) for _ in calls]


def run_all():
    for (a, k), o in zip(calls, outs):
        k = dict(k, out=o, want_mask=False)
        orig(*a, **k)


reps = 10
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    run_all()
torch.cuda.synchronize()
ref = [orig(*a, **k)[0] for a, k in calls]
torch.cuda.synchronize()
for r, o in zip(ref, outs):
    assert torch.equal(r, o)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(reps):
        run_all()
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
best, tot = 1e9, 0.0
iters = 20
for _ in range(iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e3 / (reps * len(calls))
    best, tot = min(best, t), tot + t
print(f"K1 {mode} variant={os.environ.get('TC_SAMPLE_VARIANT', '0')}: {tot / iters:.2f} us per launch (best {best:.2f}), "
      f"{len(calls)} layers x {reps} reps per graph")
