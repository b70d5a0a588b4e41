"""Stage-level parity of the bf16x3 (split-bf16) tensor-core mode (``-m gpu``): every kernel that produces or consumes
a ``TC_BF16X2`` / ``TC_F16`` operand against an fp64 evaluation of the same fp32 inputs.

bf16x3 carries 16 mantissa bits per operand (hi + lo bf16 halves) and evaluates hi*hi + lo*hi + hi*lo with fp32
accumulation, so a K = 256 product is within ~1e-5 of fp32 arithmetic; the tolerances below are the fp32 ones of
BASELINE.json's north_star relaxed only by that operand rounding (4e-5 abs + 2e-5 rel on O(1)-O(5) outputs).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

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


def test_cast_split_is_16_bit_exact(ops):
    x = rnd((333, 192), 1, 3.0)
    s = ops.cast_split(x)
    assert s.t.shape == (333, 384) and s.shape == (333, 192)
    hi = x.bfloat16()
    assert torch.equal(s.t[:, :192], hi)
    assert torch.equal(s.t[:, 192:], (x - hi.float()).bfloat16())
    # hi + lo reproduces x to 2^-17 relative
    assert ((s.float() - x).abs() <= x.abs() * 2.0 ** -16).all()


@pytest.mark.parametrize("M,N,K", [(900, 256, 256), (77, 512, 256), (333, 10, 256), (900, 24, 256), (64, 768, 256),
                                   (1, 256, 512), (130, 128, 64), (7200, 1536, 256), (1500, 128, 64)])
def test_linear_x3_plain(ops, M, N, K):
    A, W, b = rnd((M, K), 1), rnd((N, K), 2, K ** -0.5), rnd((N,), 3, 0.1)
    want = (A.double() @ W.double().t() + b.double()).relu().float()
    o32, o16 = ops.linear(ops.cast_split(A), ops.cast_split(W), b, relu=True, want_f32=True, want_bf16=N % 16 == 0,
                          out16="split")
    torch.testing.assert_close(o32, want, rtol=2e-5, atol=4e-5)
    if o16 is not None:
        assert o16.t.shape == (M, 2 * N)
        # the split output is the 16-bit rounding of the fp32 output, bit for bit
        assert torch.equal(o16.t[:, :N], o32.bfloat16())
        assert torch.equal(o16.t[:, N:], (o32 - o32.bfloat16().float()).bfloat16())


def test_linear_x3_full_epilogue_and_chaining(ops):
    M, N, K, period = 1800, 256, 512, 900
    A, W, b = rnd((M, K), 4), rnd((N, K), 5, K ** -0.5), rnd((N,), 6, 0.1)
    rb, res, res2, post = rnd((period, N), 7, 0.3), rnd((M, N), 8), rnd((M, N), 9), rnd((M, N), 10)
    gate = (torch.arange(M, device=dev()) % 3 != 0).to(torch.uint8)
    ln = (1 + 0.1 * rnd((N,), 11), 0.1 * rnd((N,), 12))
    o32, o16 = ops.linear(ops.cast_split(A), ops.cast_split(W), b, row_bias=rb, row_bias_period=period, row_gate=gate,
                          residual=res, residual2=res2, ln=ln, relu=True, post_add=post, want_f32=True, want_bf16=True,
                          out16="split")
    y = A.double() @ W.double().t() + b.double() + rb.double()[torch.arange(M, device=dev()) % period]
    y = y * gate.double().unsqueeze(1) + res.double() + res2.double()
    y = F.layer_norm(y, (N,), ln[0].double(), ln[1].double(), 1e-5).relu() + post.double()
    torch.testing.assert_close(o32, y.float(), rtol=2e-5, atol=4e-5)
    # a second Linear consuming the split output directly (activation hand-off between GEMMs)
    W2 = rnd((128, N), 13, N ** -0.5)
    z32, _ = ops.linear(o16, ops.cast_split(W2), None)
    torch.testing.assert_close(z32, (y @ W2.double().t()).float(), rtol=2e-5, atol=6e-5)
    # SIMT fallback reads split operands too (odd row pitch is not TMA-able: K = 72)
    A3, W3 = rnd((50, 72), 14), rnd((24, 72), 15, 0.1)
    s32, _ = ops.linear(ops.cast_split(A3), ops.cast_split(W3), None)
    torch.testing.assert_close(s32, (A3.double() @ W3.double().t()).float(), rtol=2e-5, atol=2e-5)


def test_linear_f16_output_saturates(ops):
    M, N, K = 256, 64, 64
    A, W = rnd((M, K), 1), rnd((N, K), 2, K ** -0.5)
    A[0] *= 1e5                                   # row 0 overflows fp16
    o32, o16 = ops.linear(ops.cast_split(A), ops.cast_split(W), None, want_f32=True, want_bf16=True, out16="f16")
    assert o16.dtype == torch.float16 and torch.isfinite(o16).all()
    assert torch.equal(o16[1:], o32[1:].half())
    assert (o16[0].abs().float() <= 65504).all() and (o16[0].abs() == 65504).any()


@pytest.mark.parametrize("Lq,Lk", [(900, 900), (130, 77), (1, 1500)])
def test_attention_f16_split_out(ops, Lq, Lk):
    """Dense attention core on fp16 operands with a split-bf16 output: vs fp64 softmax attention on the same fp16 values."""
    B, heads, E = 2, 8, 256
    q, k, v = rnd((B, Lq, E), 1).half(), rnd((B, Lk, E), 2).half(), rnd((B, Lk, E), 3).half()
    out, _ = ops.attention(q, k, v, heads, out_dtype="split")
    D = E // heads
    qh, kh, vh = (t.double().view(B, -1, heads, D).transpose(1, 2) for t in (q, k, v))
    p = torch.softmax(qh @ kh.transpose(-1, -2) / D ** 0.5, -1)
    want = (p @ vh).transpose(1, 2).reshape(B, Lq, E).float()
    assert out.t.shape == (B, Lq, 2 * E)
    # P is rounded to fp16 (2^-11) before the PV product: 5e-4 relative on O(1) outputs
    torch.testing.assert_close(out.float(), want, rtol=1e-3, atol=5e-4)
    hi = out.t[..., :E]
    assert torch.equal(out.t[..., E:], (out.float() - hi.float()).bfloat16())


def test_sparse_attention_f32_in_split_out(ops):
    from test_gpu_stages import _geometry_inputs, _oracle_mha_core        # same masked oracle as the bf16 / fp32 tests
    B, Q, R, heads, E = 2, 900, 1500, 8, 256
    radar_xy, centre, code = _geometry_inputs(B, Q, R, seed=23)
    geom = ops.radar_geometry(centre.view(B * Q, -1), code.view(B * Q, -1), synthetic.PC_RANGE, 1.0, 2.0, True)
    q, k, v = rnd((B, Q, E), 1), rnd((B, R, E), 2), rnd((B, R, E), 3)
    ref32, any32 = ops.attention(q, k, v, heads, geom=geom, key_xy=radar_xy, want_row_any=True, algo="sparse")
    out, any16 = ops.attention(q, k, v, heads, geom=geom, key_xy=radar_xy, want_row_any=True, out_dtype="split")
    assert torch.equal(any32, any16)
    assert torch.equal(out.t[..., :E], ref32.bfloat16())
    assert torch.equal(out.t[..., E:], (ref32 - ref32.bfloat16().float()).bfloat16())


def test_sample_and_point_embed_split_out(ops):
    from test_gpu_stages import _sampling_case
    feats, metas, ref, logits = _sampling_case(2, 300, "tiny", seed=6, smooth=False)
    l2i = torch.tensor(np.asarray([m["lidar2img"] for m in metas]), dtype=torch.float32, device=dev())
    for dt in (torch.bfloat16, torch.float32):
        cl = [f.to(dev()).to(dt).permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3) for f in feats]
        o32, m32 = ops.sample_fwd(cl, ref.to(dev()), l2i, logits.to(dev()), synthetic.PC_RANGE, 1600, 928, want_mask=True)
        osp, msp = ops.sample_fwd(cl, ref.to(dev()), l2i, logits.to(dev()), synthetic.PC_RANGE, 1600, 928,
                                  out_dtype="split", want_mask=True)
        assert torch.equal(m32, msp)
        assert torch.equal(osp.t[..., :256], o32.bfloat16())
        assert torch.equal(osp.t[..., 256:], (o32 - o32.bfloat16().float()).bfloat16())
    sd = {k: v.to(dev()) for k, v in synthetic.make_state_dict(1, 128).items()}
    p = "transformer.decoder.layers.1.attentions.1.position_encoder"
    x = torch.rand((700, 3), generator=torch.Generator().manual_seed(3)).to(dev())
    args = (x, sd[p + ".0.weight"], sd[p + ".0.bias"], sd[p + ".1.weight"], sd[p + ".1.bias"])
    e32, _ = ops.point_embed(*args, logit_input=True)
    _, esp = ops.point_embed(*args, logit_input=True, want_f32=False, want_bf16=True, out16="split")
    assert torch.equal(esp.t[:, :256], e32.bfloat16())
    assert torch.equal(esp.t[:, 256:], (e32 - e32.bfloat16().float()).bfloat16())
