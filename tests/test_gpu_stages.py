"""Stage-level parity of every CUDA entry point against the oracle on identical inputs (``-m gpu``).

All calls go through the C ABI (``transcar_b200.ops`` -> ctypes -> ``libtranscar_b200.so``).
Tolerances (BASELINE.json north_star): masks bit-exact; fp32 values 1e-5; bf16 values 1e-3 abs + 1e-2 rel.
The oracle runs twice where arithmetic order matters: on the GPU (same ATen kernels the reference would
use in deployment) and on the CPU (the committed golden vectors come from there).
"""
import math
import warnings

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import fusion_decoder as O
from transcar_b200 import synthetic

pytestmark = pytest.mark.gpu

BF16_ATOL, BF16_RTOL = 1e-3, 1e-2


def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops():
    from transcar_b200 import _lib, ops as _ops
    lib = _lib.load()
    assert lib.tc_check_device() == 0, lib.tc_last_error_string()
    return _ops


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev())


# ------------------------------------------------------------------------------------------- K1
def _sampling_case(B, Q, config, seed, smooth):
    feats = synthetic.make_feats(seed, B, config, smooth=smooth)
    metas = synthetic.make_img_metas(B, seed=seed)
    g = torch.Generator().manual_seed(seed + 11)
    ref = torch.rand((B, Q, 3), generator=g)
    ref[:, : Q // 8] = ref[:, : Q // 8] * 1.2 - 0.1          # some points outside [0,1] / behind cameras
    logits = torch.randn((B, Q, 24), generator=g)
    return feats, metas, ref, logits


def _oracle_sampled_sum(feats, metas, ref, logits, device):
    """T:365-373 with explicit attention logits (the Linear is tested separately)."""
    feats = [f.to(device) for f in feats]
    ref, logits = ref.to(device), logits.to(device)
    B, Q, _ = ref.shape
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, sampled, mask = O.feature_sampling(feats, ref, metas)
    sampled = torch.nan_to_num(sampled, nan=0.0)
    w = logits.view(B, 1, Q, 6, 1, 4).sigmoid() * mask
    out = (sampled * w).sum(-1).sum(-1).sum(-1).permute(0, 2, 1)
    return out, mask[:, 0, :, :, 0, 0]


@pytest.mark.parametrize("config,B,Q", [("tiny", 2, 300), ("res101", 1, 900), ("vovnet", 1, 900)])
def test_sample_fp32_vs_oracle(ops, config, B, Q):
    feats, metas, ref, logits = _sampling_case(B, Q, config, seed=5, smooth=False)
    l2i = torch.tensor(np.asarray([m["lidar2img"] for m in metas]), dtype=torch.float32, device=dev())
    cl = [ops.to_channels_last(f.to(dev())) for f in feats]
    out, mask = ops.sample_fwd(cl, ref.to(dev()), l2i, logits.to(dev()), synthetic.PC_RANGE, 1600, 928, want_mask=True)
    torch.cuda.synchronize()
    # (1) oracle on the GPU: the arithmetic the reference has in deployment
    want_g, mask_g = _oracle_sampled_sum(feats, metas, ref, logits, dev())
    assert torch.equal(mask.bool(), mask_g), "camera validity mask must be bit-exact (GPU oracle)"
    assert mask.sum().item() > 0
    torch.testing.assert_close(out, want_g, rtol=0, atol=1e-5)
    # (2) oracle on the CPU: ATen's CPU grid_sample un-normalises coordinates with a different rounding
    # (SURVEY H2), so white-noise texels allow 1e-4 here; mask still bit-exact
    want_c, mask_c = _oracle_sampled_sum(feats, metas, ref, logits, "cpu")
    assert torch.equal(mask.bool().cpu(), mask_c), "camera validity mask must be bit-exact (CPU oracle)"
    torch.testing.assert_close(out.cpu(), want_c, rtol=0, atol=5e-4)


def test_sample_smooth_feats_vs_cpu_and_gpu_oracle(ops):
    feats, metas, ref, logits = _sampling_case(2, 256, "tiny", seed=9, smooth=True)
    l2i = torch.tensor(np.asarray([m["lidar2img"] for m in metas]), dtype=torch.float32, device=dev())
    cl = [ops.to_channels_last(f.to(dev())) for f in feats]
    out, _ = ops.sample_fwd(cl, ref.to(dev()), l2i, logits.to(dev()), synthetic.PC_RANGE, 1600, 928)
    want_c, _ = _oracle_sampled_sum(feats, metas, ref, logits, "cpu")
    # the kernel follows ATen's CUDA arithmetic (x * (1/1600), CUDA un-normalisation); ATen's CPU kernels round
    # these two steps differently, which shows as ~1e-5 on band-limited features
    torch.testing.assert_close(out.cpu(), want_c, rtol=0, atol=3e-5)
    want_g, _ = _oracle_sampled_sum(feats, metas, ref, logits, dev())
    torch.testing.assert_close(out, want_g, rtol=0, atol=1e-5)


def test_sample_bf16_and_layouts(ops):
    feats, metas, ref, logits = _sampling_case(2, 300, "tiny", seed=6, smooth=False)
    l2i = torch.tensor(np.asarray([m["lidar2img"] for m in metas]), dtype=torch.float32, device=dev())
    f16 = [f.to(dev()).to(torch.bfloat16) for f in feats]
    cl16 = [f.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3) for f in f16]       # zero-copy layout
    out16, mask16 = ops.sample_fwd(cl16, ref.to(dev()), l2i, logits.to(dev()), synthetic.PC_RANGE, 1600, 928,
                                   out_dtype=torch.bfloat16, want_mask=True)
    # reference on the SAME bf16-rounded texels in fp32 arithmetic
    want, mask = _oracle_sampled_sum([f.float() for f in f16], metas, ref, logits, dev())
    assert torch.equal(mask16.bool(), mask)
    torch.testing.assert_close(out16.float(), want, rtol=BF16_RTOL, atol=BF16_ATOL)
    # fp32 output from bf16 features is exact up to accumulation order
    out32, _ = ops.sample_fwd(cl16, ref.to(dev()), l2i, logits.to(dev()), synthetic.PC_RANGE, 1600, 928)
    torch.testing.assert_close(out32, want, rtol=0, atol=2e-5)
    # NCHW fp32 hand-off goes through tc_nchw_to_nhwc (also to bf16)
    conv = ops.to_channels_last(feats[0].to(dev()), torch.bfloat16)
    assert torch.equal(conv, feats[0].to(dev()).to(torch.bfloat16))
    assert ops.is_channels_last_5d(conv)


@pytest.mark.parametrize("B,Q", [(3, 37), (5, 901), (1, 1), (7, 149)])
def test_sample_ragged_query_counts_bf16(ops, B, Q):
    """Query counts that do not divide the per-CTA query ranges or the warp count: every (sample, query) row is handed out
    exactly once by the in-kernel scheduler and lands in its own batch row (mask bit-exact, values at the bf16 bar)."""
    feats, metas, ref, logits = _sampling_case(B, Q, "tiny", seed=3 + Q % 7, smooth=False)
    l2i = torch.tensor(np.asarray([m["lidar2img"] for m in metas]), dtype=torch.float32, device=dev())
    f16 = [f.to(dev()).to(torch.bfloat16) for f in feats]
    cl16 = [f.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3) for f in f16]
    out = torch.full((B, Q, 256), float("nan"), device=dev(), dtype=torch.bfloat16)       # every row must be written
    out16, mask16 = ops.sample_fwd(cl16, ref.to(dev()), l2i, logits.to(dev()), synthetic.PC_RANGE, 1600, 928,
                                   out_dtype=torch.bfloat16, want_mask=True, out=out)
    want, mask = _oracle_sampled_sum([f.float() for f in f16], metas, ref, logits, dev())
    assert torch.equal(mask16.bool(), mask)
    assert torch.isfinite(out16.float()).all()
    torch.testing.assert_close(out16.float(), want, rtol=BF16_RTOL, atol=BF16_ATOL)


def test_sample_edge_cases(ops):
    feats, metas, _, _ = _sampling_case(1, 8, "tiny", seed=2, smooth=False)
    l2i = torch.tensor(np.asarray([m["lidar2img"] for m in metas]), dtype=torch.float32, device=dev())
    cl = [ops.to_channels_last(f.to(dev())) for f in feats]
    # empty query set
    out, mask = ops.sample_fwd(cl, torch.zeros((1, 0, 3), device=dev()), l2i, torch.zeros((1, 0, 24), device=dev()),
                               synthetic.PC_RANGE, 1600, 928, want_mask=True)
    assert out.shape == (1, 0, 256) and mask.shape == (1, 0, 6)
    # points that no camera sees -> exact zeros, all-false mask (NaN/inf-free: z clamped at 1e-5)
    ref = torch.tensor([[[0.5, 0.5, 40.0], [0.5, 0.5, -40.0], [1e9, 0.0, 0.5], [float("nan"), 0.5, 0.5]]], device=dev())
    out, mask = ops.sample_fwd(cl, ref, l2i, torch.zeros((1, 4, 24), device=dev()), synthetic.PC_RANGE, 1600, 928,
                               want_mask=True)
    want, mask_o = _oracle_sampled_sum(feats, metas, ref.cpu(), torch.zeros((1, 4, 24)), dev())
    assert torch.equal(mask.bool(), mask_o)
    assert torch.isfinite(out).all()
    torch.testing.assert_close(out, want, rtol=0, atol=1e-5)
    # non channels-last input is refused loudly
    with pytest.raises(RuntimeError):
        ops.sample_fwd([f.to(dev()) for f in feats], ref, l2i, torch.zeros((1, 4, 24), device=dev()),
                       synthetic.PC_RANGE, 1600, 928)


# ------------------------------------------------------------------------------------------- K3
def _ref_linear(A, W, bias=None, row_bias=None, period=0, gate=None, res=None, res2=None, ln=None, relu=False,
                post=None):
    y = A.double() @ W.double().t()
    if bias is not None:
        y = y + bias.double()
    if row_bias is not None:
        idx = torch.arange(A.shape[0], device=A.device) % period
        y = y + row_bias.double()[idx]
    if gate is not None:
        y = y * gate.double().unsqueeze(1)
    if res is not None:
        y = y + res.double()
    if res2 is not None:
        y = y + res2.double()
    if ln is not None:
        y = F.layer_norm(y, (y.shape[1],), ln[0].double(), ln[1].double(), 1e-5)
    if relu:
        y = y.relu()
    if post is not None:
        y = y + post.double()
    return y.float()


@pytest.mark.parametrize("M,N,K", [(900, 256, 256), (77, 512, 256), (1500, 64, 36), (333, 10, 256), (900, 24, 256),
                                   (64, 768, 256), (1, 256, 512), (130, 128, 64)])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_linear_plain(ops, M, N, K, mode):
    A, W, b = rnd((M, K), 1), rnd((N, K), 2, K ** -0.5), rnd((N,), 3, 0.1)
    if mode == "bf16":
        A16, W16 = A.bfloat16(), W.bfloat16()
        o32, o16 = ops.linear(A16, W16, b, relu=True, want_f32=True, want_bf16=True)
        want = _ref_linear(A16.float(), W16.float(), b, relu=True)
        torch.testing.assert_close(o32, want, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(o16.float(), want, rtol=BF16_RTOL, atol=BF16_ATOL)
    else:
        o32, _ = ops.linear(A, W, b, relu=True)
        torch.testing.assert_close(o32, _ref_linear(A, W, b, relu=True), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_linear_full_epilogue(ops, mode):
    M, N, K, period = 1800, 256, 512, 900
    A, W, b = rnd((M, K), 4), rnd((N, K), 5, K ** -0.5), rnd((N,), 6, 0.1)
    rb, res, res2, post = rnd((period, N), 7, 0.3), rnd((M, N), 8), rnd((M, N), 9), rnd((M, N), 10)
    gate = (torch.arange(M, device=dev()) % 3 != 0).to(torch.uint8)
    ln = (1 + 0.1 * rnd((N,), 11), 0.1 * rnd((N,), 12))
    kw = dict(row_bias=rb, row_bias_period=period, row_gate=gate, residual=res, residual2=res2, ln=ln, relu=True,
              post_add=post)
    if mode == "bf16":
        A, W = A.bfloat16(), W.bfloat16()
    o32, o16 = ops.linear(A, W, b, want_f32=True, want_bf16=True, **kw)
    want = _ref_linear(A.float(), W.float(), b, rb, period, gate, res, res2, ln, True, post)
    tol = dict(rtol=1e-5, atol=2e-5) if mode == "fp32" else dict(rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(o32, want, **tol)
    torch.testing.assert_close(o16.float(), want, rtol=BF16_RTOL, atol=BF16_ATOL * 4)
    # strided output / operand views (qkv slices, stacked outputs)
    big = torch.zeros((M, 3 * N), device=dev())
    ops.linear(A, W, b, out_f32=big[:, N:2 * N])
    torch.testing.assert_close(big[:, N:2 * N], _ref_linear(A.float(), W.float(), b), rtol=1e-4, atol=1e-4)
    assert big[:, :N].abs().sum() == 0 and big[:, 2 * N:].abs().sum() == 0


def test_linear_layernorm_nan_rows_stay_local_and_fast(ops):
    """The LayerNorm statistics travel between the CTAs of a cluster as (sum, sum of squares) pairs with a NaN-pattern
    'empty' marker: rows whose sums ARE NaN (NaN residual) must neither be mistaken for the marker (a bounded spin of
    ~40 ms per row block) nor disturb the other rows."""
    import time
    M, N, K = 7200, 256, 256
    A, W, b = rnd((M, K), 4).bfloat16(), rnd((N, K), 5, K ** -0.5).bfloat16(), rnd((N,), 6, 0.1)
    res = rnd((M, N), 8)
    ln = (1 + 0.1 * rnd((N,), 11), 0.1 * rnd((N,), 12))
    clean, _ = ops.linear(A, W, b, residual=res, ln=ln)
    bad = res.clone()
    bad[5::97, 3] = float("nan")
    ops.linear(A, W, b, residual=bad, ln=ln)                 # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, _ = ops.linear(A, W, b, residual=bad, ln=ln)
    torch.cuda.synchronize()
    assert time.perf_counter() - t0 < 0.02, "NaN statistics were taken for the empty marker (bounded spin)"
    nan_rows = torch.zeros(M, dtype=torch.bool, device=dev())
    nan_rows[5::97] = True
    assert torch.isnan(out[nan_rows]).all()
    assert torch.equal(out[~nan_rows], clean[~nan_rows])


def test_linear_matches_oracle_ffn_block(ops):
    """FFN + norm of a decoder layer, oracle weights (fp32 path, 1e-5)."""
    sd = {k: v.to(dev()) for k, v in synthetic.make_state_dict(1, 128).items()}
    p = "transformer.decoder.layers.2"
    x = rnd((640, 256), 21)
    h, _ = ops.linear(x, sd[p + ".ffns.0.layers.0.0.weight"], sd[p + ".ffns.0.layers.0.0.bias"], relu=True)
    y, _ = ops.linear(h, sd[p + ".ffns.0.layers.1.weight"], sd[p + ".ffns.0.layers.1.bias"], residual=x,
                      ln=(sd[p + ".norms.2.weight"], sd[p + ".norms.2.bias"]))
    hh = F.relu(O.lin(sd, p + ".ffns.0.layers.0.0", x))
    want = O.lnorm(sd, p + ".norms.2", x + O.lin(sd, p + ".ffns.0.layers.1", hh))
    torch.testing.assert_close(y, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("M", [7200, 640, 100])
def test_ffn_fused_matches_oracle_and_unfused(ops, M):
    """tc_ffn (one launch: both GEMMs of the feed-forward block, residual, LayerNorm) vs the oracle's FFN + norm in fp32
    and vs the two-launch bf16x3 path it replaces (mmcv FFN + norms.2; H:583-586).  M = 7200 is the benched size (57 row
    blocks, the last with 32 valid rows), 100 a single partial block."""
    sd = {k: v.to(dev()) for k, v in synthetic.make_state_dict(1, 128).items()}
    p = "transformer.decoder.layers.2"
    x = rnd((M, 256), 21 + M)
    x[3] *= 30.0                          # a row with a large dynamic range
    W1, W2 = ops.mark_static(ops.cast_split(sd[p + ".ffns.0.layers.0.0.weight"])), ops.mark_static(ops.cast_split(sd[p + ".ffns.0.layers.1.weight"]))
    b1, b2 = sd[p + ".ffns.0.layers.0.0.bias"], sd[p + ".ffns.0.layers.1.bias"]
    ln = (sd[p + ".norms.2.weight"], sd[p + ".norms.2.bias"])
    x16 = ops.cast_split(x)
    assert ops.ffn_supported(x16, W1, W2)
    y32, y16 = ops.ffn(x16, W1, b1, W2, b2, x, ln)
    torch.cuda.synchronize()
    hh = F.relu(O.lin(sd, p + ".ffns.0.layers.0.0", x))
    want = O.lnorm(sd, p + ".norms.2", x + O.lin(sd, p + ".ffns.0.layers.1", hh))
    torch.testing.assert_close(y32, want, rtol=2e-5, atol=2e-5)
    # the split 16-bit copy carries the same values to ~16 mantissa bits
    hi, lo = y16.t[:, :256].float(), y16.t[:, 256:].float()
    torch.testing.assert_close(hi + lo, y32, rtol=2e-5, atol=1e-6)
    assert torch.equal(hi, y32.bfloat16().float())
    # the two launches it replaces
    h, h16 = ops.linear(x16, W1, b1, relu=True, want_f32=False, want_bf16=True, out16="split")
    u32, _ = ops.linear(h16, W2, b2, residual=x, ln=ln, want_bf16=True, out16="split")
    torch.testing.assert_close(y32, u32, rtol=1e-5, atol=1e-5)
    # outputs are optional one by one
    only32, none16 = ops.ffn(x16, W1, b1, W2, b2, x, ln, want_16=False)
    none32, only16 = ops.ffn(x16, W1, b1, W2, b2, x, ln, want_f32=False)
    assert none16 is None and none32 is None
    assert torch.equal(only32, y32) and torch.equal(only16.t, y16.t)


def test_ffn_bad_arguments(ops):
    x = ops.cast_split(rnd((64, 256), 1))
    W1, W2 = ops.cast_split(rnd((512, 256), 2)), ops.cast_split(rnd((256, 512), 3))
    b1, b2, g = rnd((512,), 4), rnd((256,), 5), rnd((256,), 6)
    with pytest.raises(RuntimeError, match="C = 256, H = 512"):
        ops.ffn(x, ops.cast_split(rnd((256, 256), 7)), b1, W2, b2, rnd((64, 256), 8), (g, g))
    with pytest.raises(RuntimeError, match="residual has shape"):
        ops.ffn(x, W1, b1, W2, b2, rnd((64, 128), 8), (g, g))


@pytest.mark.parametrize("M", [7200, 200])
def test_mlp_fused_heads_match_oracle_and_unfused(ops, M):
    """tc_mlp (one launch: three Linear layers of a head, hidden activations on chip) vs the oracle in fp32 and vs the three
    bf16x3 tc_linear launches it replaces, tails included: the refinement branch (reg_branches, T:190-203), a regression
    head with the box tail (final_reg, H:588-600) and a classification head with its LayerNorms (final_cls, H:128-139)."""
    sd = {k: v.to(dev()) for k, v in synthetic.make_state_dict(1, 128).items()}
    x = rnd((M, 256), 77 + M)
    x16 = ops.cast_split(x)
    pc = synthetic.PC_RANGE
    ref = torch.rand((M, 3), generator=torch.Generator().manual_seed(5)).to(dev())
    W = lambda k: ops.mark_static(ops.cast_split(sd[k + ".weight"]))

    def unfused(keys, lns, tail):
        a, a16 = ops.linear(x16, W(keys[0]), sd[keys[0] + ".bias"], ln=lns[0], relu=True, want_bf16=True, out16="split")
        b, b16 = ops.linear(a16, W(keys[1]), sd[keys[1] + ".bias"], ln=lns[1], relu=True, want_bf16=True, out16="split")
        y, _ = ops.linear(b16, W(keys[2]), sd[keys[2] + ".bias"], tail=tail)
        return y

    # ---- refinement branch + reference update (+ first radar geometry)
    keys = tuple(f"reg_branches.3.{i}" for i in (0, 2, 4))
    t_f = dict(kind="ref_update", ref=ref, pc_range=pc, geom=(1.0, 2.0))
    y = ops.mlp(x16, W(keys[0]), sd[keys[0] + ".bias"], W(keys[1]), sd[keys[1] + ".bias"], W(keys[2]), sd[keys[2] + ".bias"], tail=t_f)
    want = O.lin(sd, keys[2], F.relu(O.lin(sd, keys[1], F.relu(O.lin(sd, keys[0], x)))))
    torch.testing.assert_close(y, want, rtol=2e-5, atol=2e-5)
    t_u = dict(kind="ref_update", ref=ref, pc_range=pc, geom=(1.0, 2.0))
    torch.testing.assert_close(y, unfused(keys, (None, None), t_u), rtol=1e-5, atol=1e-5)
    # the tail is a function of the row's outputs: bit-identical to the stand-alone kernels run on this result
    assert torch.equal(t_f["ref_out"], ops.ref_update(y, ref))
    assert torch.equal(t_f["geom_out"], ops.radar_geometry(t_f["ref_out"], y, pc, 1.0, 2.0, centre_is_normalised=True))

    # ---- regression head + box anchor tail into a strided destination
    keys = tuple(f"final_reg2.{i}" for i in (0, 2, 4))
    anchor = rnd((M, 10), 7, 20.0)
    out = torch.empty((3, M, 10), device=dev())[1]
    plain = ops.mlp(x16, W(keys[0]), sd[keys[0] + ".bias"], W(keys[1]), sd[keys[1] + ".bias"], W(keys[2]), sd[keys[2] + ".bias"])
    t_f = dict(kind="box", anchor=anchor, xy_col=0, z_col=4, from_norm=False, pc_range=pc, geom=(0.5, 1.0))
    ops.mlp(x16, W(keys[0]), sd[keys[0] + ".bias"], W(keys[1]), sd[keys[1] + ".bias"], W(keys[2]), sd[keys[2] + ".bias"],
            out_f32=out, tail=t_f)
    want = O.lin(sd, keys[2], F.relu(O.lin(sd, keys[1], F.relu(O.lin(sd, keys[0], x)))))
    torch.testing.assert_close(plain, want, rtol=2e-5, atol=2e-5)
    boxed = ops.box_anchor_add(plain.clone(), anchor, 0, 4, False, pc)
    assert torch.equal(out, boxed)
    assert torch.equal(t_f["geom_out"], ops.radar_geometry(boxed, boxed, pc, 0.5, 1.0, centre_is_normalised=False))

    # ---- classification head: Linear + LayerNorm + ReLU twice, then Linear
    keys = tuple(f"final_cls.{i}" for i in (0, 3, 6))
    lns = tuple((sd[f"final_cls.{i}.weight"], sd[f"final_cls.{i}.bias"]) for i in (1, 4))
    y = ops.mlp(x16, W(keys[0]), sd[keys[0] + ".bias"], W(keys[1]), sd[keys[1] + ".bias"], W(keys[2]), sd[keys[2] + ".bias"],
                ln1=lns[0], ln2=lns[1])
    h1 = F.relu(O.lnorm(sd, "final_cls.1", O.lin(sd, keys[0], x)))
    h2 = F.relu(O.lnorm(sd, "final_cls.4", O.lin(sd, keys[1], h1)))
    # two LayerNorms in the chain amplify the ~1e-5 error of a bf16x3 product: 5e-5 against the fp32 oracle, while the
    # fused and the three-launch path (same arithmetic, different summation order of the LayerNorm statistics) agree to 1e-5
    torch.testing.assert_close(y, O.lin(sd, keys[2], h2), rtol=5e-5, atol=5e-5)
    torch.testing.assert_close(y, unfused(keys, lns, None), rtol=1e-5, atol=1e-5)


def test_mlp_bad_arguments(ops):
    x = ops.cast_split(rnd((64, 256), 1))
    W1, W3 = ops.cast_split(rnd((256, 256), 2)), ops.cast_split(rnd((10, 256), 3))
    b, b3 = rnd((256,), 4), rnd((10,), 5)
    with pytest.raises(RuntimeError, match="C = 256 and N3 <= 32"):
        ops.mlp(x, W1, b, W1, b, ops.cast_split(rnd((64, 256), 6)), rnd((64,), 7))
    with pytest.raises(RuntimeError, match="out_f32 has shape"):
        ops.mlp(x, W1, b, W1, b, W3, b3, out_f32=torch.empty((64, 12), device=dev()))


@pytest.mark.parametrize("M,N,K,relu", [(12000, 1536, 256, False), (2100, 1024, 128, True)])
def test_linear_wide_tile_variant(ops, M, N, K, relu):
    """Plain large-N bf16x3 products with an fp32 output take the 128 x 256-tile kernel (csrc/linear_wide_tc.cu: the stacked
    radar K / V projection, H:578 / H:646 / H:704): vs fp32 matmul, including a partial last row block and a strided output."""
    A, W, b = rnd((M, K), 31), rnd((N, K), 32, K ** -0.5), rnd((N,), 33, 0.1)
    out = torch.zeros((M, N + 64), device=dev())[:, :N]                   # row pitch larger than N
    y, _ = ops.linear(ops.cast_split(A), ops.mark_static(ops.cast_split(W)), b, relu=relu, out_f32=out)
    want = A @ W.t() + b
    if relu:
        want = want.relu()
    # bf16x3 keeps ~16 mantissa bits per operand: 18 M outputs of magnitude up to 5 -> the worst of them is 2.6e-5 off
    torch.testing.assert_close(y, want, rtol=2e-5, atol=5e-5)
    assert y.data_ptr() == out.data_ptr()
    if relu:
        return
    # per-query row bias + fp16 / bf16 output (the self-attention in-projection's epilogue), no fp32 copy
    rb = rnd((900, N), 34, 0.5)
    want = A @ W.t() + rb.repeat((M + 899) // 900, 1)[:M]
    for out16, dt in (("f16", torch.float16), ("bf16", torch.bfloat16)):
        y32, y16 = ops.linear(ops.cast_split(A), ops.mark_static(ops.cast_split(W)), None, row_bias=rb, row_bias_period=900,
                              want_f32=False, want_bf16=True, out16=out16)
        assert y32 is None and y16.dtype == dt
        torch.testing.assert_close(y16.float(), want.to(dt).float(), rtol=2 ** -9 if dt == torch.float16 else 2 ** -6, atol=1e-3)


def test_linear_bad_arguments(ops):
    A, W = rnd((4, 8), 1), rnd((3, 8), 2)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.linear(A.cpu(), W)
    with pytest.raises(RuntimeError, match="LayerNorm needs N"):
        ops.linear(rnd((4, 8), 1), rnd((300, 8), 2), ln=(rnd((300,), 3), rnd((300,), 4)))
    o, _ = ops.linear(torch.zeros((0, 8), device=dev()), W)        # empty M is fine
    assert o.shape == (0, 3)


def test_point_embed(ops):
    sd = {k: v.to(dev()) for k, v in synthetic.make_state_dict(1, 128).items()}
    p = "transformer.decoder.layers.1.attentions.1.position_encoder"
    g = torch.Generator().manual_seed(3)
    ref = torch.rand((1000, 3), generator=g).to(dev())
    ref[:5] = torch.tensor([0.0, 1.0, 1e-7, 0.5, 1 - 1e-7], device=dev()).unsqueeze(1)
    o32, o16 = ops.point_embed(ref, sd[p + ".0.weight"], sd[p + ".0.bias"], sd[p + ".1.weight"], sd[p + ".1.bias"],
                               logit_input=True, want_bf16=True)
    want = F.relu(O.lnorm(sd, p + ".1", O.lin(sd, p + ".0", O.logit(ref))))
    torch.testing.assert_close(o32, want, rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(o16.float(), want, rtol=BF16_RTOL, atol=BF16_ATOL)
    # radar variant: raw xyz columns of a [M,36] token matrix, including the 500 padding rows
    tok = rnd((700, 36), 5, 20.0)
    tok[600:] = 500.0
    q = "radar_position_encoder"
    o32, _ = ops.point_embed(tok, sd[q + ".0.weight"], sd[q + ".0.bias"], sd[q + ".1.weight"], sd[q + ".1.bias"],
                             logit_input=False)
    want = F.relu(O.lnorm(sd, q + ".1", O.lin(sd, q + ".0", tok[:, :3])))
    torch.testing.assert_close(o32, want, rtol=1e-5, atol=2e-5)


# ------------------------------------------------------------------------------------------- K4 + mask
def _geometry_inputs(B, Q, R, seed):
    g = torch.Generator().manual_seed(seed)
    radar_xy = (torch.rand((B, R, 2), generator=g) * 102.4 - 51.2)
    radar_xy[:, R - R // 5:] = 500.0                                   # padding slots (quirk Q5)
    centre = torch.rand((B, Q, 3), generator=g)                          # normalised
    code = torch.randn((B, Q, 10), generator=g) * 0.3
    code[..., 3] += 1.0
    return radar_xy.to(dev()), centre.to(dev()), code.to(dev())


def _oracle_blocked(centre_m, code, radar_xy, lo, hi):
    out = []
    for b in range(centre_m.shape[0]):
        out.append(O.radar_block_mask(centre_m[b:b + 1, :, :2].clone(), code[b:b + 1, :, 3].exp(), -code[b:b + 1, :, 6],
                                      -code[b:b + 1, :, 7], radar_xy[b:b + 1], lo, hi))
    return torch.stack(out)


@pytest.mark.parametrize("lo,hi,normalised", [(1.0, 2.0, True), (0.5, 1.0, False)])
def test_radar_mask_bit_exact(ops, lo, hi, normalised):
    B, Q, R = 2, 900, 1500
    radar_xy, centre, code = _geometry_inputs(B, Q, R, seed=17)
    centre_m = O.to_metres(centre)
    src = centre if normalised else centre_m
    geom = ops.radar_geometry(src.view(B * Q, 3), code.view(B * Q, 10), synthetic.PC_RANGE, lo, hi, normalised)
    blocked, row_any = ops.radar_mask(geom, radar_xy, B, Q, R)
    want_gpu = _oracle_blocked(centre_m, code, radar_xy, lo, hi)                       # torch.cdist on the GPU
    want_cpu = _oracle_blocked(centre_m.cpu(), code.cpu(), radar_xy.cpu(), lo, hi)     # and on the CPU
    n_allowed = int((~want_gpu).sum())
    assert n_allowed > 100
    assert torch.equal(blocked.bool(), want_gpu), \
        f"radar mask differs from torch.cdist(GPU) in {(blocked.bool() != want_gpu).sum().item()} of {want_gpu.numel()} bits"
    assert torch.equal(blocked.bool().cpu(), want_cpu), \
        f"radar mask differs from torch.cdist(CPU) in {(blocked.bool().cpu() != want_cpu).sum().item()} bits"
    assert torch.equal(row_any.bool(), (~want_gpu).any(-1))


def _oracle_mha_core(q, k, v, heads, blocked=None):
    """softmax(q k^T / sqrt(d) + mask) v per head, fp64; rows with no allowed key -> 0."""
    B, Lq, E = q.shape
    D = E // heads
    qh = q.double().view(B, Lq, heads, D).transpose(1, 2)
    kh = k.double().view(B, -1, heads, D).transpose(1, 2)
    vh = v.double().view(B, -1, heads, D).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(D)
    if blocked is not None:
        s = s.masked_fill(blocked.unsqueeze(1), float("-inf"))
    p = torch.softmax(s, dim=-1)
    p = torch.nan_to_num(p, nan=0.0)
    return (p @ vh).transpose(1, 2).reshape(B, Lq, E).float()


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("Lq,Lk", [(900, 900), (130, 77), (1, 1500)])
def test_attention_dense(ops, mode, Lq, Lk):
    B, heads, E = 2, 8, 256
    q, k, v = rnd((B, Lq, E), 1), rnd((B, Lk, E), 2), rnd((B, Lk, E), 3)
    if mode == "bf16":
        q, k, v = q.bfloat16(), k.bfloat16(), v.bfloat16()
    out, _ = ops.attention(q, k, v, heads)
    want = _oracle_mha_core(q.float(), k.float(), v.float(), heads)
    if mode == "fp32":
        torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-5)
    else:
        torch.testing.assert_close(out.float(), want, rtol=BF16_RTOL, atol=BF16_ATOL * 2)


@pytest.mark.parametrize("mode,algo", [("fp32", "sparse"), ("fp32", "simt"), ("bf16", "sparse"), ("bf16", "tensor"),
                                       ("bf16", "simt"), ("bf16", "auto")])
def test_attention_radar_mask_in_kernel(ops, mode, algo):
    B, Q, R, heads, E = 2, 900, 1500, 8, 256
    radar_xy, centre, code = _geometry_inputs(B, Q, R, seed=23)
    geom = ops.radar_geometry(centre.view(B * Q, 3), code.view(B * Q, 10), synthetic.PC_RANGE, 1.0, 2.0, True)
    blocked = _oracle_blocked(O.to_metres(centre), code, radar_xy, 1.0, 2.0)
    # qkv as strided views of one packed buffer (how the engine hands them over)
    qbuf, kvbuf = rnd((B, Q, E), 4), rnd((B, R, 2 * E), 5)
    if mode == "bf16":
        qbuf, kvbuf = qbuf.bfloat16(), kvbuf.bfloat16()
    out, row_any = ops.attention(qbuf, kvbuf[:, :, :E], kvbuf[:, :, E:], heads, geom=geom, key_xy=radar_xy,
                                 want_row_any=True, algo=algo)
    want = _oracle_mha_core(qbuf.float(), kvbuf[:, :, :E].float(), kvbuf[:, :, E:].float(), heads, blocked)
    any_allowed = (~blocked).any(-1)
    assert torch.equal(row_any.bool(), any_allowed)
    assert 0 < any_allowed.sum() < B * Q
    assert (out[~any_allowed] == 0).all(), "rows without an allowed key must produce exact zeros (quirk Q6)"
    if mode == "fp32":
        torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-5)
    elif algo == "tensor":
        # dense-tile tcgen05 kernel: P is rounded to bf16 before P.V (both MMA operands must share one 16-bit format),
        # which costs up to 2^-9 * sum|p v| on rows that attend to two or three keys with cancelling values
        torch.testing.assert_close(out.float(), want, rtol=BF16_RTOL, atol=BF16_ATOL * 6)
    else:
        torch.testing.assert_close(out.float(), want, rtol=BF16_RTOL, atol=BF16_ATOL * 2)


@pytest.mark.parametrize("algo", ["sparse", "simt"])
def test_attention_nan_radar_point_is_never_attended(ops, algo):
    """torch.cdist clamps the squared distance with clamp_min (NaN stays NaN), so a NaN radar coordinate is blocked for
    every query (H:549-571); fmaxf(NaN, 0) = 0 would have made it the nearest point of all of them."""
    B, Q, R, heads, E = 1, 300, 64, 8, 256
    radar_xy, centre, code = _geometry_inputs(B, Q, R, seed=37)
    radar_xy[0, 5, 0] = float("nan")
    geom = ops.radar_geometry(centre.view(B * Q, 3), code.view(B * Q, 10), synthetic.PC_RANGE, 1.0, 2.0, True)
    blocked = _oracle_blocked(O.to_metres(centre), code, radar_xy, 1.0, 2.0)
    assert blocked[0, :, 5].all()
    q, kv = rnd((B, Q, E), 4), rnd((B, R, 2 * E), 5)
    out, row_any = ops.attention(q, kv[:, :, :E], kv[:, :, E:], heads, geom=geom, key_xy=radar_xy, want_row_any=True, algo=algo)
    assert torch.equal(row_any.bool(), (~blocked).any(-1))
    want = _oracle_mha_core(q, kv[:, :, :E], kv[:, :, E:], heads, blocked)
    torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-5)
    mask, _ = ops.radar_mask(geom, radar_xy, B, Q, R)
    assert torch.equal(mask.bool(), blocked)


def test_attention_matches_nn_multiheadattention(ops):
    """Whole nn.MultiheadAttention (in-proj, core, out-proj) with a bool mask == the oracle's mha(), fp32."""
    sd = {k: v.to(dev()) for k, v in synthetic.make_state_dict(1, 128).items()}
    B, Q, R, E = 1, 300, 500, 256
    radar_xy, centre, code = _geometry_inputs(B, Q, R, seed=29)
    blocked = _oracle_blocked(O.to_metres(centre), code, radar_xy, 1.0, 2.0)[0]
    rows = torch.where((~blocked).any(1))[0]
    x, kv = rnd((Q, 1, E), 6), rnd((R, 1, E), 7)
    want = O.mha(sd, "rf_multihead_attn", x[rows], kv, kv, attn_mask=blocked[rows])
    w, b = sd["rf_multihead_attn.in_proj_weight"], sd["rf_multihead_attn.in_proj_bias"]
    qp, _ = ops.linear(x[:, 0], w[:E], b[:E])
    kvp, _ = ops.linear(kv[:, 0], w[E:], b[E:])
    geom = ops.radar_geometry(centre.view(Q, 3), code.view(Q, 10), synthetic.PC_RANGE, 1.0, 2.0, True)
    att, row_any = ops.attention(qp.view(1, Q, E), kvp.view(1, R, 2 * E)[:, :, :E], kvp.view(1, R, 2 * E)[:, :, E:], 8,
                                 geom=geom, key_xy=radar_xy, want_row_any=True)
    y, _ = ops.linear(att.view(Q, E), sd["rf_multihead_attn.out_proj.weight"], sd["rf_multihead_attn.out_proj.bias"],
                      row_gate=row_any.view(Q))
    torch.testing.assert_close(y[rows], want[:, 0], rtol=1e-5, atol=1e-5)
    mask = torch.ones(Q, dtype=torch.bool, device=dev())
    mask[rows] = False
    assert (y[mask] == 0).all()


# ------------------------------------------------------------------------------------------- pointwise + decode
def test_ref_update_and_anchor(ops):
    g = torch.Generator().manual_seed(31)
    ref = torch.rand((1200, 3), generator=g).to(dev())
    ref[:4, 0] = torch.tensor([0.0, 1.0, 1e-6, 1 - 1e-6], device=dev())
    code = rnd((1200, 10), 32)
    new = ops.ref_update(code, ref)
    want = torch.zeros_like(ref)
    want[..., :2] = code[..., :2] + O.logit(ref[..., :2])
    want[..., 2:3] = code[..., 4:5] + O.logit(ref[..., 2:3])
    torch.testing.assert_close(new, want.sigmoid(), rtol=1e-6, atol=1e-6)
    # layer 1 anchor: normalised ref -> metres for x,y; z added as is (quirk Q3)
    reg = rnd((1200, 10), 33)
    want = reg.clone()
    m = O.to_metres(ref)
    want[:, 0:2] += m[:, 0:2]
    want[:, 4] += ref[:, 2]
    got = ops.box_anchor_add(reg.clone(), ref, 0, 2, True, synthetic.PC_RANGE)
    assert torch.equal(got, want)
    # later layers: previous code columns 0,1,4
    prev = rnd((1200, 10), 34, 10.0)
    want = reg.clone()
    want[:, 0:2] += prev[:, 0:2]
    want[:, 4] += prev[:, 4]
    assert torch.equal(ops.box_anchor_add(reg.clone(), prev, 0, 4, False, synthetic.PC_RANGE), want)


def test_decode_vs_oracle(ops):
    B, Q = 3, 900
    cls, code = rnd((B, Q, 10), 41, 2.0), rnd((B, Q, 10), 42)
    code[..., 0:2] *= 40.0          # some centres outside the +-61.2 post range
    rng = [-61.2, -61.2, -10.0, 61.2, 61.2, 10.0]
    boxes, scores, labels, keep = ops.decode(cls, code, 300, rng)
    for b in range(B):
        want = O.nms_free_decode(cls[b], code[b], 300, 10, rng)
        k = keep[b].bool()
        assert 0 < k.sum() < 300
        torch.testing.assert_close(scores[b][k], want["scores"], rtol=0, atol=1e-6)
        assert torch.equal(labels[b][k].long(), want["labels"])
        torch.testing.assert_close(boxes[b][k], want["bboxes"], rtol=1e-6, atol=1e-5)


def test_decode_ties_records_and_short_inputs(ops):
    """Radix-select decode: exact ties at the selection threshold are broken by the lowest flat index (deterministic),
    the single-tensor record output equals the four-tensor one, and fewer scores than max_num are zero-padded."""
    rng = [-61.2, -61.2, -10.0, 61.2, 61.2, 10.0]
    B, Q = 2, 900
    cls = torch.zeros((B, Q, 10), device=dev())                      # every score = 0.5: 9000-way tie
    cls[0, 5, 3] = 2.0
    cls[0, 700, 9] = 1.0
    cls[1, :, 4] = -1.0                                               # sample 1: class 4 loses against the rest
    code = rnd((B, Q, 10), 43)
    boxes, scores, labels, keep = ops.decode(cls, code, 300, rng)
    flat0 = torch.cat([torch.tensor([5 * 10 + 3, 700 * 10 + 9]), torch.tensor([i for i in range(400) if i not in (53,)][:298])])
    assert torch.equal(labels[0].cpu().long(), flat0 % 10)
    torch.testing.assert_close(scores[0, :3].cpu(), torch.tensor([2.0, 1.0, 0.0]).sigmoid(), rtol=0, atol=1e-7)
    want1 = torch.tensor([i for i in range(400) if i % 10 != 4][:300])
    assert torch.equal(labels[1].cpu().long(), want1 % 10)
    torch.testing.assert_close(boxes[1, :, 0].cpu(), code[1, want1 // 10, 0].cpu(), rtol=0, atol=0)
    rec = ops.decode(cls, code, 300, rng, records=True)
    assert rec.shape == (B, 300, 12)
    assert torch.equal(rec[..., :9], boxes) and torch.equal(rec[..., 9], scores)
    assert torch.equal(rec[..., 10], labels.float()) and torch.equal(rec[..., 11], keep.float())
    # fewer scores than max_num
    b2, s2, l2, k2 = ops.decode(cls[:, :20].contiguous(), code[:, :20].contiguous(), 300, rng)
    assert (k2[:, 200:] == 0).all() and (s2[:, 200:] == 0).all() and (b2[:, 200:] == 0).all()
    assert (s2[:, :200] > 0).all()
    # random scores once more against torch.topk (values; ties have measure zero)
    cls3 = rnd((3, Q, 10), 44, 3.0)
    _, s3, l3, _ = ops.decode(cls3, rnd((3, Q, 10), 45), 300, rng)
    top, idx = cls3.sigmoid().view(3, -1).topk(300, dim=1)
    torch.testing.assert_close(s3, top, rtol=0, atol=1e-7)
    assert torch.equal(l3.long(), idx % 10)


@pytest.mark.parametrize("mode", ["fp32", "bf16", "bf16x3"])
def test_linear_fused_tails_match_the_standalone_kernels(ops, mode):
    """tc_linear tails (reference update T:195-203; box anchor H:596-600 / :664-665 + next-layer mask geometry H:615-635)
    produce bit-identical results to tc_ref_update / tc_box_anchor_add / tc_radar_geometry run after a plain tc_linear."""
    M, K = 1800, 256
    A, W, b = rnd((M, K), 1, 0.5), rnd((10, K), 2, K ** -0.5), rnd((10,), 3, 0.1)
    if mode == "bf16":
        A, W = A.bfloat16(), W.bfloat16()
    elif mode == "bf16x3":
        A, W = ops.cast_split(A), ops.cast_split(W)
    ref = torch.rand((M, 3), generator=torch.Generator().manual_seed(5)).to(dev())
    pc = synthetic.PC_RANGE
    plain, _ = ops.linear(A, W, b)
    # ---- reference update (+ geometry of the first radar layer)
    tail = dict(kind="ref_update", ref=ref, pc_range=pc, geom=(1.0, 2.0))
    y, _ = ops.linear(A, W, b, tail=tail)
    assert torch.equal(y, plain)
    want_ref = ops.ref_update(plain, ref)
    assert torch.equal(tail["ref_out"], want_ref)
    assert torch.equal(tail["geom_out"], ops.radar_geometry(want_ref, plain, pc, 1.0, 2.0, centre_is_normalised=True))
    # ---- box anchor from normalised reference points (layer 1) and from a previous code (layers 2, 3)
    for anchor, xy, z, norm, clamp in ((ref, 0, 2, True, (1.0, 2.0)), (rnd((M, 10), 7, 20.0), 0, 4, False, (0.5, 1.0))):
        tail = dict(kind="box", anchor=anchor, xy_col=xy, z_col=z, from_norm=norm, pc_range=pc, geom=clamp)
        out = torch.empty((3, M, 10), device=dev())[1]              # a strided destination like reg_all[li]
        ops.linear(A, W, b, out_f32=out, tail=tail)
        want = ops.box_anchor_add(plain.clone(), anchor, xy, z, norm, pc)
        assert torch.equal(out, want)
        assert torch.equal(tail["geom_out"], ops.radar_geometry(want, want, pc, *clamp, centre_is_normalised=False))
    with pytest.raises(RuntimeError, match="tail needs"):
        ops.linear(A, W, b, relu=True, tail=dict(kind="ref_update", ref=ref, pc_range=pc))
