"""Sparse radar attention on the bench configuration's own geometry (radar layer masks of one synthetic step): attended rows,
allowed keys per row, and the device time of the launch (`reps` launches in one CUDA graph).  Usage: sparse_attn_bench.py [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcar_b200 import ops, plugin, synthetic
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B, Q = 8, 900
cfg = synthetic.head_config(Q); cfg["precision"] = "bf16x3"
head = plugin.build_head(cfg); head.load_state_dict(synthetic.make_state_dict(0, Q)); head = head.cuda().eval()
feats = [f.to(torch.bfloat16).cuda().permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3) for f in synthetic.make_feats(0, B, "res101", smooth=True)]
with torch.no_grad():
    out = head(feats, synthetic.make_img_metas(B, seed=0), return_aux=True)
aux = out["aux"]
key_xy = aux["key_xy"]
R = key_xy.shape[1]
q = torch.randn(B, Q, 256, device="cuda"); kv = torch.randn(B, R, 512, device="cuda")
for li in range(3):
    blocked, _ = ops.radar_mask(aux[f"radar{li}.geom"], key_xy, B, Q, R)
    n = (blocked == 0).sum(-1).float()
    res = {}
    for name in ("scan",):
        def body():
            for _ in range(reps):
                ops.attention(q, kv[:, :, :256], kv[:, :, 256:], 8, geom=aux[f"radar{li}.geom"], key_xy=key_xy, want_row_any=True, algo="sparse")
        body(); torch.cuda.synchronize()
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s): body()
        torch.cuda.current_stream().wait_stream(s)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr): body()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
        res[name] = best
    print(f"radar layer {li}: rows with a key {int(aux[f'radar{li}.row_any'].sum())} of {B * Q}; allowed keys per attended row: mean "
          f"{n[n > 0].mean():.1f}, max {int(n.max())}; launch {res['scan']:.2f} us")
