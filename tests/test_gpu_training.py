"""Training variant (BASELINE.json configs[4], reference recipe tools/train.py:238-252): gradients of every radar-head
parameter from the library's backward kernels vs PyTorch autograd through the oracle on identical inputs (``-m gpu``).

Upstream gradients are random tensors (the Hungarian loss stays in PyTorch and is out of scope); parity is on
d loss / d parameter for loss = <all_cls_scores, Gc> + <all_bbox_preds, Gr>."""
import warnings

import numpy as np
import pytest
import torch

from oracle import fusion_decoder as O
from transcar_b200 import synthetic

pytestmark = pytest.mark.gpu


def _setup(Q, B, seed, levels="tiny"):
    from transcar_b200 import plugin
    sd = synthetic.make_state_dict(seed=seed, num_query=Q)
    cfg = synthetic.head_config(num_query=Q)
    cfg["precision"] = "fp32"
    head = plugin.build_head(cfg)
    head.load_state_dict(sd, strict=True)
    head = head.cuda().train()
    feats = synthetic.make_feats(seed, B, levels)
    metas = synthetic.make_img_metas(B, seed=seed)
    return sd, head, feats, metas


def _oracle_grads(sd, feats, metas, Gc, Gr, names, device):
    sd = {k: v.to(device).clone().requires_grad_(k in names) for k, v in sd.items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = O.head_forward(sd, [f.to(device) for f in feats], metas)
    loss = (out["all_cls_scores"] * Gc.to(device)).sum() + (out["all_bbox_preds"] * Gr.to(device)).sum()
    loss.backward()
    return out, {k: sd[k].grad for k in names}


def test_radar_head_gradients_vs_oracle_autograd():
    from transcar_b200 import _lib
    from transcar_b200.training import trainable_names
    Q, B, seed = 128, 2, 7
    sd, head, feats, metas = _setup(Q, B, seed)
    names = set(trainable_names(sd.keys()))
    for k, p in head.named_parameters():                      # reference recipe: only the radar head trains
        p.requires_grad_(k in names)
    g = torch.Generator().manual_seed(99)
    Gc = torch.randn((3, B, Q, 10), generator=g)
    Gr = torch.randn((3, B, Q, 10), generator=g)
    n0 = _lib.launch_count()
    out = head([f.cuda() for f in feats], metas)              # training mode + grad enabled -> forward_train
    loss = (out["all_cls_scores"] * Gc.cuda()).sum() + (out["all_bbox_preds"] * Gr.cuda()).sum()
    loss.backward()
    torch.cuda.synchronize()
    assert _lib.launch_count() - n0 > 300, "backward must run library kernels"
    want_out, want = _oracle_grads(sd, feats, metas, Gc, Gr, names, "cuda")
    # forward of the training path == inference path == oracle
    torch.testing.assert_close(out["all_cls_scores"], want_out["all_cls_scores"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(out["all_bbox_preds"], want_out["all_bbox_preds"], rtol=1e-4, atol=1e-4)
    got = dict(head.named_parameters())
    checked = 0
    for k in sorted(names):
        gk, wk = got[k].grad, want[k]
        assert gk is not None, f"no gradient for {k}"
        if wk is None:
            wk = torch.zeros_like(gk)
        scale = max(float(wk.abs().max()), 1e-3)
        err = float((gk - wk).abs().max())
        assert err <= 2e-3 * scale + 1e-5, f"{k}: max |dgrad| {err:.3e} vs scale {scale:.3e}"
        checked += 1
    assert checked == len(names) and checked >= 60
    # frozen parameters stay without gradient
    assert got["query_embedding.weight"].grad is None
    # the attention path actually carried gradient
    assert float(got["rf_multihead_attn.in_proj_weight"].grad[128 * 0:256].abs().sum()) > 0
    assert float(got["radar_feat_encoder.0.weight"].grad.abs().sum()) > 0


def test_training_step_reduces_loss_and_bucket_matches():
    """A few SGD steps on a fixed target through ``GradBucket`` (single process): loss goes down."""
    from transcar_b200 import sharding
    from transcar_b200.training import trainable_names
    Q, B, seed = 128, 2, 11
    sd, head, feats, metas = _setup(Q, B, seed)
    names = set(trainable_names(sd.keys()))
    for k, p in head.named_parameters():
        p.requires_grad_(k in names)
    params = [p for p in head.parameters() if p.requires_grad]
    bucket = sharding.GradBucket(params, n_scalars=6)
    feats_c = [f.cuda() for f in feats]
    g = torch.Generator().manual_seed(5)
    target = torch.randn((3, B, Q, 10), generator=g).cuda()
    losses = []
    for step in range(4):
        bucket.zero()
        out = head(feats_c, metas)
        loss = ((out["all_bbox_preds"] - target) ** 2).mean() + (out["all_cls_scores"] ** 2).mean()
        loss.backward()
        bucket.all_reduce()                                   # world size 1: no-op, exercises the call path
        with torch.no_grad():
            for p in params:
                p.add_(p.grad, alpha=-2e-3)
        losses.append(float(loss.detach()))
    assert np.isfinite(losses).all()
    assert losses[-1] < losses[0], losses


def test_forward_train_is_stream_safe_and_keeps_the_decoder_engine():
    """(a) the decoder's last refinement runs on a side stream - forward_train must read it after a join:
    identical outputs with and without parallel branches, repeatedly; (b) optimizer steps on the radar head must not
    rebuild the (frozen) decoder engine."""
    from transcar_b200.training import trainable_names
    Q, B, seed = 128, 2, 13
    sd, head, feats, metas = _setup(Q, B, seed)
    names = set(trainable_names(sd.keys()))
    for k, p in head.named_parameters():
        p.requires_grad_(k in names)
    feats_c = [f.cuda() for f in feats]
    eng = head.decoder_engine()
    eng.use_branches = False
    ref_out = head(feats_c, metas)
    eng.use_branches = True
    for _ in range(5):
        out = head(feats_c, metas)
        assert torch.equal(out["all_cls_scores"], ref_out["all_cls_scores"])
        assert torch.equal(out["all_bbox_preds"], ref_out["all_bbox_preds"])
    (out["all_cls_scores"].sum() + out["all_bbox_preds"].sum()).backward()
    with torch.no_grad():
        for p in head.parameters():
            if p.requires_grad:
                p.add_(p.grad, alpha=-1e-3)               # bumps the parameter version like optimizer.step()
    head(feats_c, metas)
    assert head.decoder_engine() is eng, "radar-head updates must not rebuild the decoder engine"
    # un-frozen decoder parameters are refused, not silently left without gradient
    head.transformer.reference_points.weight.requires_grad_(True)
    with pytest.raises(NotImplementedError, match="only the radar head trains"):
        head(feats_c, metas)
