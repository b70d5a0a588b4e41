"""ORACLE - test infrastructure only, never a product path.

CPU restatement (plain PyTorch eager, fp32, device-agnostic) of the TransCAR fusion-decoder hot
path, written as pure functions over a reference-named state dict.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import
this package; ``transcar_b200`` never does and fails loudly when its CUDA library is missing.

Parity status: PINNED against the reference itself.  ``oracle/gen_golden.py`` imports the
reference modules verbatim from ``/root/reference`` (through ``oracle/refstubs.py``), runs them on
the seeded synthetic inputs of ``transcar_b200/synthetic.py`` and stores their outputs in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays this restatement against those
vectors.  (The reference ships no tests or golden vectors of its own for this path - SURVEY.md F6.)
The decoder-layer wrapper (mmcv ``DetrTransformerDecoderLayer``/``MultiheadAttention``/``FFN``, mmcv-full
1.3.8-1.4.0, not vendored under ``/root/reference``) is restated from its published behaviour
(SURVEY.md appendix A); the same restatement is what ``refstubs.py`` hands the reference files.

Reference files (all under ``/root/reference/projects/mmdet3d_plugin/``):
  T = ``models/utils/detr3d_transformer.py``   H = ``models/dense_heads/detr3d_head.py``
  C = ``core/bbox/coders/nms_free_coder.py``   U = ``core/bbox/util.py``
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

PC_RANGE = (-51.2, -51.2, -5.0, 51.2, 51.2, 3.0)
NUM_HEADS = 8
RADAR_SLOTS = 1500
RADAR_PAD = 500.0
# (lo, hi) clamp of the attention radius per radar layer - H:567, H:635, H:693 (quirk Q7)
RADIUS_CLAMP = ((1.0, 2.0), (1.0, 2.0), (0.5, 1.0))


# Training-mode dropout of the reference (p = 0.1 at every site listed in transcar_b200/training.py).  The oracle does not
# draw random numbers: a test installs ``DROPOUT = f(site, tensor, **ctx) -> tensor`` that multiplies by the SAME mask the
# library regenerates from (seed, stream), so that gradients can be compared element by element.  None = eval mode.
DROPOUT = None


def _drop(site, t, **ctx):
    return t if DROPOUT is None else DROPOUT(site, t, **ctx)


# ------------------------------------------------------------------ small helpers
def lin(sd, key, x):
    return F.linear(x, sd[key + ".weight"], sd[key + ".bias"])


def lnorm(sd, key, x):
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], 1e-5)


def logit(x, eps=1e-5):
    """T:17-32 ``inverse_sigmoid``."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def mha(sd, prefix, q, k, v, attn_mask=None, site=None, **ctx):
    """``nn.MultiheadAttention(256, 8)`` slow path (``need_weights=True`` default: baddbmm ->
    softmax -> bmm, bool mask turned into -inf), eval mode.  q/k/v are ``[L,B,E]``.  H:578, and the
    ``.attn`` member of mmcv's wrapper.  With a ``DROPOUT`` hook installed the same steps are spelled out so that the
    attention-probability dropout (``dropout=0.1`` of the module) can be applied with a given mask."""
    if DROPOUT is not None:
        E, H = q.shape[-1], NUM_HEADS
        D = E // H
        w, b = sd[prefix + ".in_proj_weight"], sd[prefix + ".in_proj_bias"]
        Lq, B, _ = q.shape
        Lk = k.shape[0]
        qh = (F.linear(q, w[:E], b[:E]) * D ** -0.5).reshape(Lq, B * H, D).transpose(0, 1)
        kh = F.linear(k, w[E:2 * E], b[E:2 * E]).reshape(Lk, B * H, D).transpose(0, 1)
        vh = F.linear(v, w[2 * E:], b[2 * E:]).reshape(Lk, B * H, D).transpose(0, 1)
        s = qh @ kh.transpose(1, 2)
        if attn_mask is not None:
            s = s.masked_fill(attn_mask.unsqueeze(0), float("-inf"))
        pr = _drop(site + ".probs", s.softmax(-1), **ctx)                       # [B*H, Lq, Lk]
        o = (pr @ vh).transpose(0, 1).reshape(Lq, B, E)
        return F.linear(o, sd[prefix + ".out_proj.weight"], sd[prefix + ".out_proj.bias"])
    out, _ = F.multi_head_attention_forward(
        q, k, v, q.shape[-1], NUM_HEADS,
        sd[prefix + ".in_proj_weight"], sd[prefix + ".in_proj_bias"],
        None, None, False, 0.0,
        sd[prefix + ".out_proj.weight"], sd[prefix + ".out_proj.bias"],
        training=False, key_padding_mask=None, need_weights=True, attn_mask=attn_mask)
    return out


def to_metres(ref, pc=PC_RANGE):
    out = ref.clone()
    out[..., 0:1] = ref[..., 0:1] * (pc[3] - pc[0]) + pc[0]
    out[..., 1:2] = ref[..., 1:2] * (pc[4] - pc[1]) + pc[1]
    out[..., 2:3] = ref[..., 2:3] * (pc[5] - pc[2]) + pc[2]
    return out


# ------------------------------------------------------------------ a6: feature_sampling
def project_points(ref, img_metas, pc=PC_RANGE):
    """T:382-411.  ``ref [B,Q,3]`` in [0,1] -> grid ``[B,N,Q,2]`` in [-1,1] coordinates and the
    camera-validity mask ``[B,N,Q]`` (bool).  Image W/H come from sample 0 (quirk Q2)."""
    l2i = np.asarray([m["lidar2img"] for m in img_metas])
    l2i = ref.new_tensor(l2i)                                   # float64 -> fp32, (B,N,4,4)
    pts = to_metres(ref, pc)
    pts = torch.cat((pts, torch.ones_like(pts[..., :1])), -1)   # homogeneous
    B, Q = pts.shape[:2]
    N = l2i.size(1)
    pts = pts.view(B, 1, Q, 4).repeat(1, N, 1, 1).unsqueeze(-1)
    mats = l2i.view(B, N, 1, 4, 4).repeat(1, 1, Q, 1, 1)
    cam = torch.matmul(mats, pts).squeeze(-1)                   # [B,N,Q,4]
    eps = 1e-5
    valid = cam[..., 2:3] > eps
    uv = cam[..., 0:2] / torch.maximum(cam[..., 2:3], torch.ones_like(cam[..., 2:3]) * eps)
    uv[..., 0] /= img_metas[0]["img_shape"][0][1]
    uv[..., 1] /= img_metas[0]["img_shape"][0][0]
    grid = (uv - 0.5) * 2
    valid = (valid & (grid[..., 0:1] > -1.0) & (grid[..., 0:1] < 1.0)
             & (grid[..., 1:2] > -1.0) & (grid[..., 1:2] < 1.0))
    return grid, valid.squeeze(-1)


def feature_sampling(mlvl_feats, ref, img_metas, pc=PC_RANGE):
    """T:381-422.  Returns ``(ref3d [B,Q,3], sampled [B,C,Q,N,1,L], mask [B,1,Q,N,1,1] bool)``."""
    grid, valid = project_points(ref, img_metas, pc)
    B, N, Q = valid.shape
    mask = valid.view(B, N, 1, Q, 1, 1).permute(0, 2, 3, 1, 4, 5)
    per_level = []
    for feat in mlvl_feats:
        b, n, c, h, w = feat.shape
        s = F.grid_sample(feat.reshape(b * n, c, h, w), grid.view(b * n, Q, 1, 2),
                          mode="bilinear", padding_mode="zeros", align_corners=False)
        per_level.append(s.view(b, n, c, Q, 1).permute(0, 2, 3, 1, 4))
    sampled = torch.stack(per_level, -1).view(B, mlvl_feats[0].shape[2], Q, N, 1, len(mlvl_feats))
    return ref.clone(), sampled, mask


def sampled_sum(sd, prefix, q_plus_pos, mlvl_feats, ref, img_metas):
    """T:362-373: the masked, sigmoid-weighted sum over (cam, point, level).
    ``q_plus_pos [B,Q,C]`` -> ``(summed [B,Q,C], mask [B,1,Q,N,1,1])``.  This is what the fused
    sampling kernel K1 produces."""
    B, Q, _ = q_plus_pos.shape
    n_levels = len(mlvl_feats)
    n_cams = mlvl_feats[0].shape[1]
    # num_points (1 in the TransCAR configs, 5 by default) is whatever the Linear's width implies: T:362-365
    aw = lin(sd, prefix + ".attention_weights", q_plus_pos).view(B, 1, Q, n_cams, -1, n_levels)
    ref3d, out, mask = feature_sampling(mlvl_feats, ref, img_metas)
    out = torch.nan_to_num(out, nan=0.0, posinf=float("inf"), neginf=float("-inf"))
    w = aw.sigmoid() * mask
    out = (out * w).sum(-1).sum(-1).sum(-1)                     # [B,C,Q]
    return out.permute(0, 2, 1), mask


# ------------------------------------------------------------------ a5/a7: Detr3DCrossAtten
def cross_atten(sd, prefix, query, query_pos, mlvl_feats, ref, img_metas):
    """T:302-378 in eval mode.  ``query, query_pos [Q,B,C]`` -> ``[Q,B,C]``.
    The residual is the query *before* the positional term is added (quirk Q1)."""
    residual = query
    q = (query + query_pos).permute(1, 0, 2)
    s, _ = sampled_sum(sd, prefix, q, mlvl_feats, ref, img_metas)
    out = lin(sd, prefix + ".output_proj", s.permute(1, 0, 2))
    p = logit(ref)
    p = F.relu(lnorm(sd, prefix + ".position_encoder.1", lin(sd, prefix + ".position_encoder.0", p)))
    p = F.relu(lnorm(sd, prefix + ".position_encoder.4", lin(sd, prefix + ".position_encoder.3", p)))
    return _drop(prefix + ".cross_out", out) + residual + p.permute(1, 0, 2)      # T:378 self.dropout(output)


# ------------------------------------------------------------------ a4: decoder layer (mmcv wrapper)
def decoder_layer(sd, prefix, query, query_pos, mlvl_feats, ref, img_metas):
    """mmcv ``DetrTransformerDecoderLayer`` with ``operation_order = (self_attn, norm, cross_attn,
    norm, ffn, norm)`` - cfg ``detr3d_res101_gridmask.py:65-82``; semantics per SURVEY appendix A."""
    x = query
    qk = x + query_pos
    x = x + _drop(prefix + ".attn_out", mha(sd, prefix + ".attentions.0.attn", qk, qk, x, site=prefix))
    x = lnorm(sd, prefix + ".norms.0", x)
    x = cross_atten(sd, prefix + ".attentions.1", x, query_pos, mlvl_feats, ref, img_metas)
    x = lnorm(sd, prefix + ".norms.1", x)
    h = _drop(prefix + ".ffn_hidden", F.relu(lin(sd, prefix + ".ffns.0.layers.0.0", x)))
    x = x + _drop(prefix + ".ffn_out", lin(sd, prefix + ".ffns.0.layers.1", h))
    x = lnorm(sd, prefix + ".norms.2", x)
    return x


def reg_branch(sd, prefix, x):
    h = F.relu(lin(sd, prefix + ".0", x))
    h = F.relu(lin(sd, prefix + ".2", h))
    return lin(sd, prefix + ".4", h)


def cls_branch(sd, prefix, x):
    h = F.relu(lnorm(sd, prefix + ".1", lin(sd, prefix + ".0", x)))
    h = F.relu(lnorm(sd, prefix + ".4", lin(sd, prefix + ".3", h)))
    return lin(sd, prefix + ".6", h)


# ------------------------------------------------------------------ a2/a3: transformer + decoder
def transformer(sd, mlvl_feats, img_metas, num_layers=6, capture=None):
    """T:75-139 + T:155-214.  Returns ``(hs [L,Q,B,C], init_ref [B,Q,3], inter_refs [L,B,Q,3])``."""
    B = mlvl_feats[0].size(0)
    emb = sd["query_embedding.weight"]
    C = emb.shape[1] // 2
    query_pos, query = torch.split(emb, C, dim=1)
    query_pos = query_pos.unsqueeze(0).expand(B, -1, -1)
    query = query.unsqueeze(0).expand(B, -1, -1)
    ref = lin(sd, "transformer.reference_points", query_pos).sigmoid()
    init_ref = ref
    x = query.permute(1, 0, 2)
    pos = query_pos.permute(1, 0, 2)
    hs, refs = [], []
    for lid in range(num_layers):
        x = decoder_layer(sd, f"transformer.decoder.layers.{lid}", x, pos, mlvl_feats, ref, img_metas)
        tmp = reg_branch(sd, f"reg_branches.{lid}", x.permute(1, 0, 2))
        new = torch.zeros_like(ref)
        new[..., :2] = tmp[..., :2] + logit(ref[..., :2])
        new[..., 2:3] = tmp[..., 4:5] + logit(ref[..., 2:3])
        ref = new.sigmoid().detach()                               # T:203: the refined points carry no gradient onwards
        hs.append(x)
        refs.append(ref)
        if capture is not None:
            capture[f"dec{lid}.out"] = x
            capture[f"dec{lid}.ref"] = ref
    return torch.stack(hs), init_ref, torch.stack(refs)


# ------------------------------------------------------------------ a9/a10: radar tokens + encoders
def pad_tokens(tokens, device, dtype=torch.float32):
    """H:523-530.  ``[n,36]`` -> ``[1,1500,36]`` with 500 in every unused slot (quirk Q5)."""
    t = torch.as_tensor(np.asarray(tokens), dtype=dtype, device=device).reshape(-1, 36)
    fill = min(RADAR_SLOTS, t.shape[0])
    out = torch.full((1, RADAR_SLOTS, 36), RADAR_PAD, dtype=dtype, device=device)
    out[0, :fill] = t[:fill]
    return out, fill


def radar_encode(sd, tokens):
    """H:531-536.  ``tokens [1,R,36]`` -> ``[1,R,C]``; padding rows are encoded like any other."""
    p = tokens[..., :3]
    p = F.relu(lnorm(sd, "radar_position_encoder.1", lin(sd, "radar_position_encoder.0", p)))
    p = F.relu(lnorm(sd, "radar_position_encoder.4", lin(sd, "radar_position_encoder.3", p)))
    f = F.relu(lin(sd, "radar_feat_encoder.0", tokens))
    f = F.relu(lin(sd, "radar_feat_encoder.2", f))
    f = F.relu(lin(sd, "radar_feat_encoder.4", f))
    return p + f


# ------------------------------------------------------------------ a11: distance mask
def radar_block_mask(centre_xy, length, rot_s, rot_c, radar_xy, lo, hi):
    """H:549-571 (and :619-640, :675-698).  ``centre_xy [1,Q,2]`` metres, ``length/rot_s/rot_c [1,Q]``,
    ``radar_xy [1,R,2]`` -> ``blocked [Q,R]`` bool (True = key not attended).
    ``torch.cdist`` keeps the reference's mm-based Euclidean formula."""
    front = centre_xy.clone()
    rear = centre_xy.clone()
    front[..., 0] = front[..., 0] + length * 0.25 * rot_s
    front[..., 1] = front[..., 1] + length * 0.25 * rot_c
    rear[..., 0] = rear[..., 0] - length * 0.25 * rot_s
    rear[..., 1] = rear[..., 1] - length * 0.25 * rot_c
    d_c = torch.cdist(centre_xy, radar_xy, p=2.0)
    d_f = torch.cdist(front, radar_xy, p=2.0)
    d_r = torch.cdist(rear, radar_xy, p=2.0)
    radii = (length / 2.0).reshape(-1, 1).repeat(1, radar_xy.shape[1])
    radii = torch.clamp(radii, min=lo, max=hi)
    return ~((d_c[0] < radii) + (d_f[0] < radii) + (d_r[0] < radii))


# ------------------------------------------------------------------ a12-a14: one radar layer
def radar_layer(sd, idx, x, kv, blocked, sample=0):
    """H:573-593 for layer ``idx`` in {0,1,2}.  ``x [Q,1,C]``, ``kv [R,1,C]``, ``blocked [Q,R]``.
    Rows with no allowed key skip attention but still go through LN/FFN/LN (quirk Q6)."""
    s = ("", "_2", "_3")[idx]
    m = ("", "2", "3")[idx]
    rows = torch.where((blocked == False).any(dim=1))[0]      # noqa: E712
    x = x.clone()
    if rows.numel() > 0:
        y = mha(sd, "rf_multihead_attn" + m, x[rows], kv, kv, attn_mask=blocked[rows], site=f"radar{idx}", rows=rows,
                sample=sample)
        x[rows] = x[rows] + _drop(f"radar{idx}.attn_out", y, rows=rows, sample=sample)         # rf_dropout2 (H:581)
    x = lnorm(sd, "rf_norm2" + s, x)
    h = _drop(f"radar{idx}.ffn_hidden", F.relu(lin(sd, "rf_linear1" + s, x)), sample=sample)   # rf_dropout (H:584)
    x = x + _drop(f"radar{idx}.ffn_out", lin(sd, "rf_linear2" + s, h), sample=sample)          # rf_dropout3 (H:585)
    x = lnorm(sd, "rf_norm3" + s, x)
    xt = x.permute(1, 0, 2)
    cls = cls_branch(sd, "final_cls" + m, xt)
    reg = reg_branch(sd, "final_reg" + m, xt)
    return x, cls, reg, rows


# ------------------------------------------------------------------ a8: whole head forward
def head_forward(sd, mlvl_feats, img_metas, num_layers=6, capture=None):
    """H:248-740 in eval mode, generalised to B>1 by looping the (batch-1 only) radar block over
    samples.  ``img_metas[b]['radar_tokens']`` replaces the devkit disk reads (H:301-521).
    Returns ``dict(all_cls_scores [3,B,Q,10], all_bbox_preds [3,B,Q,10])``."""
    hs, init_ref, inter_refs = transformer(sd, mlvl_feats, img_metas, num_layers, capture)
    hs = hs.permute(0, 2, 1, 3)                                  # [L,B,Q,C]
    last = num_layers - 1
    # H:277-298: only tmp of the last level survives (and only columns 3,6,7 of it are read).
    prev_ref = init_ref if last == 0 else inter_refs[last - 1]
    tmp_all = reg_branch(sd, f"reg_branches.{last}", hs[last])   # raw; cols 3,6,7 untouched by H:287-293
    B = hs.shape[1]
    cls_out, reg_out = [], []
    for b in range(B):
        dev = hs.device
        tokens, fill = pad_tokens(img_metas[b]["radar_tokens"], dev)
        kv = radar_encode(sd, tokens).permute(1, 0, 2)           # [R,1,C]
        radar_xy = tokens[:, :, :2]
        x = hs[last][b:b + 1].permute(1, 0, 2).clone()           # [Q,1,C]
        ref = inter_refs[-1][b:b + 1].clone()                    # [1,Q,3] normalised
        tmp = tmp_all[b:b + 1]
        cls_b, reg_b = [], []
        # --- layer 1: centre = refined ref in metres, box length/heading from reg_branches[5]
        centre = to_metres(ref)[..., :2]
        ref_xy_m = centre
        ref_z = ref[..., 2:3]                                    # stays normalised (quirk Q3)
        for li in range(3):
            length = tmp[..., 3].exp()
            rot_s = -tmp[..., 6]
            rot_c = -tmp[..., 7]
            lo, hi = RADIUS_CLAMP[li]
            blocked = radar_block_mask(centre.clone(), length, rot_s, rot_c, radar_xy, lo, hi)
            x, cls, reg, rows = radar_layer(sd, li, x, kv, blocked, sample=b)
            reg = reg.clone()
            reg[..., 0:2] = reg[..., 0:2] + ref_xy_m             # H:599, H:664, H:722
            reg[..., 4:5] = reg[..., 4:5] + ref_z                # H:600, H:665, H:723
            if capture is not None:
                capture[f"b{b}.radar{li}.blocked"] = blocked
                capture[f"b{b}.radar{li}.rows"] = rows
                capture[f"b{b}.radar{li}.x"] = x
            cls_b.append(cls)
            reg_b.append(reg)
            # next layer: centre / refs = this layer's regression (H:615-617, H:671-673)
            tmp = reg
            centre = reg[..., 0:2]
            ref_xy_m = reg[..., 0:2]
            ref_z = reg[..., 4:5]
        cls_out.append(torch.cat(cls_b, 0))                      # [3,Q,10]
        reg_out.append(torch.cat(reg_b, 0))
    return dict(all_cls_scores=torch.stack(cls_out, 1), all_bbox_preds=torch.stack(reg_out, 1),
                enc_cls_scores=None, enc_bbox_preds=None)


# ------------------------------------------------------------------ N1: NMS-free decode
def denormalize_bbox(code):
    """U:26-52 for 10-d codes: (cx,cy,w,l,cz,h,sin,cos,vx,vy) -> (cx,cy,cz,w,l,h,rot,vx,vy)."""
    rot = torch.atan2(code[..., 6:7], code[..., 7:8])
    return torch.cat([code[..., 0:1], code[..., 1:2], code[..., 4:5],
                      code[..., 2:3].exp(), code[..., 3:4].exp(), code[..., 5:6].exp(),
                      rot, code[..., 8:9], code[..., 9:10]], dim=-1)


def nms_free_decode(cls_scores, bbox_preds, max_num=300, num_classes=10,
                    post_center_range=(-61.2, -61.2, -10.0, 61.2, 61.2, 10.0)):
    """C:39-90 for one sample: sigmoid -> top-k over Q*classes -> denormalise -> centre-range filter."""
    scores, idx = cls_scores.sigmoid().view(-1).topk(max_num)
    labels = idx % num_classes
    boxes = denormalize_bbox(bbox_preds[idx // num_classes])
    rng = torch.tensor(post_center_range, device=scores.device)
    keep = (boxes[..., :3] >= rng[:3]).all(1) & (boxes[..., :3] <= rng[3:]).all(1)
    return dict(bboxes=boxes[keep], scores=scores[keep], labels=labels[keep])
