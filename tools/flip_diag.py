"""Diagnostic for the benched-config parity test: thresholded decisions that differ between the bf16x3 and fp32 engines,
with their margins (distance to the circle radius under the fp32 engine's geometry), for the fused / unfused kernel sets."""
import os, sys, warnings
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcar_b200 import ops, plugin, synthetic

Q, B, seed = 900, 8, int(os.environ.get("SEED", 0))
sd = synthetic.make_state_dict(seed=seed, num_query=Q)
feats16 = [f.to(torch.bfloat16) for f in synthetic.make_feats(seed, B, "res101", smooth=True)]
metas = synthetic.make_img_metas(B, seed=seed)
cl16 = [f.cuda().permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3) for f in feats16]
outs = {}
for precision in ("fp32", "bf16x3"):
    cfg = synthetic.head_config(num_query=Q)
    cfg["precision"] = precision
    head = plugin.build_head(cfg)
    head.load_state_dict(sd, strict=True)
    head = head.cuda().eval()
    with torch.no_grad():
        feats = cl16 if precision == "bf16x3" else [f.float() for f in cl16]
        outs[precision] = head(feats, metas, return_aux=True)
    torch.cuda.synchronize()
    del head
x3, f32 = outs["bf16x3"], outs["fp32"]
print("max |cls diff| vs fp32 engine", float((x3["all_cls_scores"] - f32["all_cls_scores"]).abs().max()),
      "median", float((x3["all_cls_scores"] - f32["all_cls_scores"]).abs().median()))
cam = [int((a != b).sum()) for a, b in zip(x3["aux"]["cam_masks"], f32["aux"]["cam_masks"])]
print("camera-mask bits differing per layer", cam)
prev_rows = set()
for li in range(3):
    g1, g2 = x3["aux"][f"radar{li}.geom"], f32["aux"][f"radar{li}.geom"]
    m1, r1 = ops.radar_mask(g1, x3["aux"]["key_xy"], B, Q, 1500)
    m2, r2 = ops.radar_mask(g2, f32["aux"]["key_xy"], B, Q, 1500)
    idx = (m1 != m2).nonzero().cpu().numpy()
    gd = (g1 - g2).abs().max().item()
    print(f"radar layer {li}: geometry max |diff| {gd:.3e}; {len(idx)} mask bits differ; rows with different row_any: {int((r1 != r2).sum())}")
    g = g2.view(B, Q, 8).double().cpu().numpy(); k = f32["aux"]["key_xy"].view(B, -1, 2).double().cpu().numpy()
    for b, q, r in idx:
        c = g[b, q]
        d = [np.hypot(*(c[2 * i:2 * i + 2] - k[b, r])) for i in range(3)]
        marg = min(abs(x - c[6]) for x in d)
        print(f"   sample {b} query {q} key {r}: margin {marg:.2e} m (radius {c[6]:.3f}); row flipped earlier: {(b, q) in prev_rows}")
    for b, q in (r1 != r2).nonzero().cpu().numpy():
        prev_rows.add((b, q))
