"""Backward kernels of the unfrozen-decoder training variant (BASELINE.json configs[4]: "grid_sample scatter + attention
bwd") against PyTorch autograd through the oracle's own formulation on identical inputs (``-m gpu``)."""
import warnings

import numpy as np
import pytest
import torch

from oracle import fusion_decoder as O
from transcar_b200 import synthetic

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops():
    from transcar_b200 import _lib, ops as _ops
    assert _lib.load().tc_check_device() == 0
    return _ops


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev())


@pytest.mark.parametrize("feat_dtype", [torch.float32, torch.bfloat16])
def test_sample_bwd_vs_autograd(ops, feat_dtype):
    """d loss / d {feature maps, attention logits, reference points} of the fused sampling step (T:367-373, 381-422)."""
    B, Q, seed = 2, 200, 5
    feats = [f.to(dev()).to(feat_dtype) for f in synthetic.make_feats(seed, B, "tiny", smooth=True)]
    metas = synthetic.make_img_metas(B, seed=seed)
    g = torch.Generator().manual_seed(seed + 11)
    ref = (torch.rand((B, Q, 3), generator=g) * 0.9 + 0.05).to(dev())
    logits = torch.randn((B, Q, 24), generator=g).to(dev())
    dout = rnd((B, Q, 256), 3)
    # ---- autograd through the oracle's formulation (fp32 arithmetic on the same texel values)
    fa = [f.float().clone().requires_grad_(True) for f in feats]
    ra, la = ref.clone().requires_grad_(True), logits.clone().requires_grad_(True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, sampled, mask = O.feature_sampling(fa, ra, metas)
    sampled = torch.nan_to_num(sampled, nan=0.0)
    w = la.view(B, 1, Q, 6, 1, 4).sigmoid() * mask
    out = (sampled * w).sum(-1).sum(-1).sum(-1).permute(0, 2, 1)
    (out * dout).sum().backward()
    # ---- library kernel
    cl = [f.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3) for f in feats]
    l2i = torch.tensor(np.asarray([m["lidar2img"] for m in metas]), dtype=torch.float32, device=dev())
    d_feats, d_logits, d_ref = ops.sample_bwd(cl, ref, l2i, logits, synthetic.PC_RANGE, 1600, 928, dout,
                                              want_feat_grad=True, want_logit_grad=True, want_ref_grad=True)
    assert mask.sum() > 0
    torch.testing.assert_close(d_logits, la.grad, rtol=1e-4, atol=1e-5)
    for got, want in zip(d_feats, fa):
        torch.testing.assert_close(got, want.grad, rtol=1e-4, atol=1e-5)
    # grid gradients are differences of neighbouring texels times large Jacobians (focal length / depth): compare at the
    # scale of the gradient itself
    scale = float(ra.grad.abs().max())
    assert scale > 0
    assert float((d_ref - ra.grad).abs().max()) <= 2e-4 * scale + 1e-6
    # accumulation into existing maps
    d2, _, _ = ops.sample_bwd(cl, ref, l2i, logits, synthetic.PC_RANGE, 1600, 928, dout, d_feats=d_feats, want_logit_grad=False,
                              want_ref_grad=True)
    torch.testing.assert_close(d2[0], 2 * fa[0].grad, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("Lq,Lk", [(300, 300), (130, 77), (900, 900)])
def test_attention_dense_bwd_vs_autograd(ops, Lq, Lk):
    B, heads, E = 2, 8, 256
    D = E // heads
    q, k, v, dout = rnd((B, Lq, E), 1), rnd((B, Lk, E), 2), rnd((B, Lk, E), 3), rnd((B, Lq, E), 4)
    qa, ka, va = (t.clone().requires_grad_(True) for t in (q, k, v))
    qh, kh, vh = (t.view(B, -1, heads, D).transpose(1, 2) for t in (qa, ka, va))
    p = torch.softmax((qh * D ** -0.5) @ kh.transpose(-1, -2), -1)
    o = (p @ vh).transpose(1, 2).reshape(B, Lq, E)
    (o * dout).sum().backward()
    # strided q/k/v views of one [B, L, 3E] buffer, as the training forward keeps them
    qkv = torch.cat([q, q, q], -1) if Lq != Lk else torch.cat([q, k, v], -1)
    kk, vv = (qkv[..., E:2 * E], qkv[..., 2 * E:]) if Lq == Lk else (k, v)
    dq, dk, dv = ops.attention_dense_bwd(qkv[..., :E], kk, vv, o.detach(), dout, heads)
    torch.testing.assert_close(dq, qa.grad, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(dk, ka.grad, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(dv, va.grad, rtol=1e-4, atol=2e-5)


def test_logit_and_sigmoid_bwd(ops):
    x = torch.rand(4096, generator=torch.Generator().manual_seed(1)).to(dev()) * 1.2 - 0.1
    x[:4] = torch.tensor([0.0, 1.0, 1e-6, 1 - 1e-7])
    g = rnd((4096,), 2)
    xa = x.clone().requires_grad_(True)
    (O.logit(xa) * g).sum().backward()
    torch.testing.assert_close(ops.logit_bwd(g, x)[4:], xa.grad[4:], rtol=1e-5, atol=1e-6)
    ya = rnd((4096,), 3).requires_grad_(True)
    y = ya.sigmoid()
    (y * g).sum().backward()
    torch.testing.assert_close(ops.sigmoid_bwd(g, y.detach()), ya.grad, rtol=1e-5, atol=1e-7)
