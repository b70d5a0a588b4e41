import os, sys, warnings
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import fusion_decoder as O
from transcar_b200 import plugin, synthetic, ops
from transcar_b200.training import trainable_names, decoder_trainable_names
Q, B, seed = 96, 2, 17
sd = synthetic.make_state_dict(seed=seed, num_query=Q)
feats = synthetic.make_feats(seed, B, "tiny", smooth=True)
metas = synthetic.make_img_metas(B, seed=seed)
g = torch.Generator().manual_seed(5)
Gc, Gr = torch.randn((3, B, Q, 10), generator=g).cuda(), torch.randn((3, B, Q, 10), generator=g).cuda()
res = {}
for precision in ("fp32", "bf16x3"):
    cfg = synthetic.head_config(num_query=Q); cfg["precision"] = precision
    head = plugin.build_head(cfg); head.load_state_dict(sd, strict=True); head = head.cuda().train()
    names = set(trainable_names(sd.keys()))
    if "unfrozen" in sys.argv: names |= set(decoder_trainable_names(sd.keys()))
    for k, p in head.named_parameters(): p.requires_grad_(k in names)
    fc = [ops.to_channels_last(f.cuda()) for f in feats]
    out = head(fc, metas)
    ((out["all_cls_scores"] * Gc).sum() + (out["all_bbox_preds"] * Gr).sum()).backward()
    torch.cuda.synchronize()
    res[precision] = ({k: p.grad.clone() for k, p in head.named_parameters() if p.grad is not None}, {k: v.detach().clone() for k, v in out.items() if v is not None},
                      {f"radar{i}.rows": None for i in range(3)})
a, b = res["fp32"][0], res["bf16x3"][0]
print("fwd max diff cls", float((res["fp32"][1]["all_cls_scores"] - res["bf16x3"][1]["all_cls_scores"]).abs().max()),
      "reg", float((res["fp32"][1]["all_bbox_preds"] - res["bf16x3"][1]["all_bbox_preds"]).abs().max()))
rows = []
for k in sorted(a):
    scale = max(float(a[k].abs().max()), 1e-3)
    rows.append((float((a[k] - b[k]).abs().max()) / scale, k, scale))
for r in sorted(rows, reverse=True)[:25]: print(f"{r[0]:.3e} {r[1]} (scale {r[2]:.3e})")
print("median rel", sorted(rows)[len(rows)//2][0])
