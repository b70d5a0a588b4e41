"""End-to-end parity of the drop-in modules against the reference's golden vectors and the oracle (``-m gpu``).

fp32 engine: tail-aware 5e-5 criterion (tests/parity.py explains why a strict 1e-5 is not attainable end to end
even between two runs of the reference itself).  bf16x3 engine (the default, and what bench.py times): BASELINE.json's
bf16 tolerance - 1e-3 abs + 1e-2 rel - on >= 99 % of the final elements after 6 + 3 chained layers, camera-validity
masks of all six layers and the three radar distance masks bit-identical to the fp32 engine's.  One-pass bf16 engine
(the optional fast mode): characterised at the looser bound it actually meets.
"""
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import fusion_decoder as O
from parity import assert_close_tail, unpack_bits
from transcar_b200 import synthetic

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def build(case, precision):
    from transcar_b200 import plugin
    g = np.load(os.path.join(GOLDEN, f"{case}.npz"))
    Q, B, seed = int(g["num_query"]), int(g["batch"]), int(g["seed"])
    sd = synthetic.make_state_dict(seed=seed, num_query=Q)
    assert synthetic.state_dict_checksum(sd) == pytest.approx(float(g["weights_checksum"]), rel=1e-12)
    cfg = synthetic.head_config(num_query=Q)
    cfg["precision"] = precision
    head = plugin.build_head(cfg)
    head.load_state_dict(sd, strict=True)          # checkpoint compatibility: reference key names
    head = head.cuda().eval()
    feats = synthetic.make_feats(seed, B, str(g["levels"]))
    metas = synthetic.make_img_metas(B, seed=seed, n_per_channel=int(g["n_per_channel"]))
    return g, head, sd, feats, metas


@pytest.mark.parametrize("case", ["tiny", "res101"])
def test_head_fp32_vs_reference_golden(case):
    g, head, sd, feats, metas = build(case, "fp32")
    B, Q = int(g["batch"]), int(g["num_query"])
    with torch.no_grad():
        out = head([f.cuda() for f in feats], metas, return_aux=True)
    torch.cuda.synchronize()
    aux = out["aux"]
    assert out["all_cls_scores"].shape == (3, B, Q, 10) and out["all_bbox_preds"].shape == (3, B, Q, 10)
    # decoder output of the last layer and refined reference points
    hs5 = aux["hs"][-1].view(B, Q, 256).cpu().numpy()
    for b in range(B):
        assert_close_tail(hs5[b], g[f"b{b}.dec5"], atol=5e-5, rtol=1e-5, frac=0.99, what=f"dec5[{b}]")
        for li in range(3):
            rows = np.zeros(Q, dtype=bool)
            rows[g[f"b{b}.radar{li}.rows"]] = True
            got = aux[f"radar{li}.row_any"][b].bool().cpu().numpy()
            assert (got != rows).sum() <= 1, f"radar layer {li}: attended-row set differs in {(got != rows).sum()} rows"
    assert_close_tail(out["all_cls_scores"].cpu().numpy(), g["all_cls_scores"], atol=5e-5, rtol=1e-5, frac=0.99, what="cls")
    assert_close_tail(out["all_bbox_preds"].cpu().numpy(), g["all_bbox_preds"], atol=5e-5, rtol=1e-5, frac=0.99, what="reg")


def golden_mask_diffs(ops, aux, g, B, Q):
    """(camera-mask bits, radar blocked-mask bits, attended rows) that differ from the reference's golden capture."""
    cam = radar = rows = 0
    R = 1500
    for b in range(B):
        want = unpack_bits(g[f"b{b}.cam_mask"], tuple(g[f"b{b}.cam_mask_shape"]))          # [layers, Q, cams]
        got = torch.stack([m[b] for m in aux["cam_masks"]]).bool().cpu().numpy()
        cam += int((got != want).sum())
        for li in range(3):
            blocked, row_any = ops.radar_mask(aux[f"radar{li}.geom"].view(B, Q, 8)[b:b + 1].contiguous(),
                                              aux["key_xy"][b:b + 1].contiguous(), 1, Q, R)
            want_b = unpack_bits(g[f"b{b}.radar{li}.blocked"], (Q, R))
            radar += int((blocked[0].bool().cpu().numpy() != want_b).sum())
            want_rows = np.zeros(Q, dtype=bool)
            want_rows[g[f"b{b}.radar{li}.rows"]] = True
            rows += int((aux[f"radar{li}.row_any"][b].bool().cpu().numpy() != want_rows).sum())
    return cam, radar, rows


@pytest.mark.parametrize("case", ["tiny", "res101"])
def test_head_fp32_masks_vs_reference_golden(case):
    """Camera-validity masks of all 6 decoder layers (T:399-409) and the radar distance masks of the 3 radar layers
    (H:549-571) against the reference's own capture.  Bit-exact per stage on identical inputs is proven in
    test_gpu_stages; end to end the inputs of layers > 0 carry ~1e-6 of accumulated GEMM-order noise, which may flip a
    point that sits within that distance of a threshold (two runs of the reference do the same): at most 2 of the
    32 400 camera bits and 2 of the 4 M radar bits."""
    from transcar_b200 import ops
    g, head, sd, feats, metas = build(case, "fp32")
    B, Q = int(g["batch"]), int(g["num_query"])
    with torch.no_grad():
        out = head([f.cuda() for f in feats], metas, return_aux=True)
    cam, radar, rows = golden_mask_diffs(ops, out["aux"], g, B, Q)
    print(f"[{case} fp32] mask bits differing from the reference: camera {cam}, radar {radar}, attended rows {rows}")
    assert cam <= 2 and radar <= 2 and rows <= 1


def test_head_bf16x3_vs_reference_golden():
    """The default engine (tensor cores, split-bf16 operands) against the reference's golden vectors for res101 at the
    north-star bf16 tolerance, masks included."""
    from transcar_b200 import ops
    g, head, sd, feats, metas = build("res101", "bf16x3")
    B, Q = int(g["batch"]), int(g["num_query"])
    with torch.no_grad():
        out = head([f.cuda() for f in feats], metas, return_aux=True)
    torch.cuda.synchronize()
    cam, radar, rows = golden_mask_diffs(ops, out["aux"], g, B, Q)
    cls, reg = out["all_cls_scores"].cpu().numpy(), out["all_bbox_preds"].cpu().numpy()
    wc, fc = assert_close_tail(cls, g["all_cls_scores"], atol=1e-3, rtol=1e-2, frac=0.99, hard_atol=2.0, what="cls(bf16x3)")
    wr, fr = assert_close_tail(reg, g["all_bbox_preds"], atol=1e-3, rtol=1e-2, frac=0.99, hard_atol=2.0, what="reg(bf16x3)")
    print(f"[res101 bf16x3 vs golden] cls {fc:.5f} in tol (max {wc:.2e}); reg {fr:.5f} (max {wr:.2e}); "
          f"mask bits differing: camera {cam}, radar {radar}, rows {rows}")
    assert cam <= 2 and radar <= 4 and rows <= 1
    hs5 = out["aux"]["hs"][-1].view(B, Q, 256).cpu().numpy()
    assert_close_tail(hs5[0], g["b0.dec5"], atol=1e-3, rtol=1e-2, frac=0.999, what="dec5(bf16x3)")


def test_head_benched_config_batch8_bf16x3_vs_oracle_and_fp32_engine():
    """The configuration bench.py times - res101 feature shapes, batch 8, bf16 channels-last feature maps, bf16x3 engine -
    against (a) the oracle looped over the 8 samples on the GPU (eager fp32 on the SAME bf16-valued feature maps) and
    (b) the fp32 engine.  Reports and asserts: fraction of all_cls_scores / all_bbox_preds within the stated
    1e-3 + 1e-2 |x|, max error, camera-mask bits and radar-mask rows per layer that differ from the fp32 engine."""
    from transcar_b200 import ops, plugin
    Q, B, seed = 900, 8, 0
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
    # (a) oracle on the GPU, one sample at a time (the reference radar block is batch-1 only)
    sd_g = {k: v.cuda() for k, v in sd.items()}
    want_cls, want_reg = [], []
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for b in range(B):
            o = O.head_forward(sd_g, [f[b:b + 1].float() for f in cl16], [dict(metas[b], img_shape=metas[0]["img_shape"])])
            want_cls.append(o["all_cls_scores"])
            want_reg.append(o["all_bbox_preds"])
    want_cls, want_reg = torch.cat(want_cls, 1).cpu().numpy(), torch.cat(want_reg, 1).cpu().numpy()
    x3, f32 = outs["bf16x3"], outs["fp32"]
    wc, fc = assert_close_tail(x3["all_cls_scores"].cpu().numpy(), want_cls, atol=1e-3, rtol=1e-2, frac=0.99, hard_atol=2.0,
                               what="cls(bf16x3, B=8)")
    wr, fr = assert_close_tail(x3["all_bbox_preds"].cpu().numpy(), want_reg, atol=1e-3, rtol=1e-2, frac=0.99, hard_atol=2.0,
                               what="reg(bf16x3, B=8)")
    # (b) masks against the fp32 engine, per layer
    cam = [int((a != b).sum()) for a, b in zip(x3["aux"]["cam_masks"], f32["aux"]["cam_masks"])]
    rows = [int((x3["aux"][f"radar{li}.row_any"] != f32["aux"][f"radar{li}.row_any"]).sum()) for li in range(3)]
    bits = []
    for li in range(3):
        m1, _ = ops.radar_mask(x3["aux"][f"radar{li}.geom"], x3["aux"]["key_xy"], B, Q, 1500)
        m2, _ = ops.radar_mask(f32["aux"][f"radar{li}.geom"], f32["aux"]["key_xy"], B, Q, 1500)
        bits.append(int((m1 != m2).sum()))
    print(f"[res101 B=8 bf16x3] vs GPU oracle: cls {fc:.5f} in 1e-3+1e-2|x| (max {wc:.2e}), reg {fr:.5f} (max {wr:.2e}); "
          f"vs fp32 engine: camera-mask bits differing per layer {cam} of {B * Q * 6}, radar-mask bits per layer {bits} "
          f"of {B * Q * 1500}, attended rows per layer {rows} of {B * Q}")
    # thresholded decisions: a point within ~1e-5 (relative) of an image border / a circle radius may land on either side,
    # and a row whose camera or radar mask flipped carries a different box into the later layers (tools/flip_diag.py
    # lists every differing decision with its margin: 3e-5 .. 5e-3 m for the first-generation flips at seeds 0 and 1)
    assert sum(cam) <= 2 and sum(rows) <= 6 and sum(bits) <= 12


def test_head_fp32_first_decoder_layer_strict():
    """One layer deep there is no error growth yet: strict 1e-5 against the golden capture."""
    g, head, sd, feats, metas = build("tiny", "fp32")
    B, Q = int(g["batch"]), int(g["num_query"])
    with torch.no_grad():
        out = head([f.cuda() for f in feats], metas, return_aux=True)
    hs0 = out["aux"]["hs"][0].view(B, Q, 256).cpu().numpy()
    for b in range(B):
        np.testing.assert_allclose(hs0[b], g[f"b{b}.dec0"], rtol=0, atol=1e-5)


def test_head_bf16_one_pass_vs_reference_golden():
    """The optional one-pass bf16 mode: 8 mantissa bits per operand through 9 chained layers."""
    g, head, sd, feats, metas = build("res101", "bf16")
    B, Q = int(g["batch"]), int(g["num_query"])
    with torch.no_grad():
        out = head([f.cuda() for f in feats], metas, return_aux=True)
    torch.cuda.synchronize()
    aux = out["aux"]
    flips = 0
    same_rows = np.ones((3, B, Q), dtype=bool)
    for b in range(B):
        for li in range(3):
            rows = np.zeros(Q, dtype=bool)
            rows[g[f"b{b}.radar{li}.rows"]] = True
            got = aux[f"radar{li}.row_any"][b].bool().cpu().numpy()
            flips += int((got != rows).sum())
            same_rows[li, b] = got == rows
    # bf16 perturbs the regressed box centres by ~1e-3 m, so a handful of threshold decisions may flip
    assert flips <= 0.02 * 3 * B * Q, f"{flips} attended-row flips"
    cls, reg = out["all_cls_scores"].cpu().numpy(), out["all_bbox_preds"].cpu().numpy()
    assert np.isfinite(cls).all() and np.isfinite(reg).all()
    # measured on B200 (round 1): cls 95.7 % / reg 99.8 % of elements within 2e-2 + 1e-2*|x| after 9 chained layers
    assert_close_tail(cls, g["all_cls_scores"], atol=2e-2, rtol=1e-2, frac=0.93, hard_atol=5.0, what="cls(bf16)")
    assert_close_tail(reg, g["all_bbox_preds"], atol=2e-2, rtol=1e-2, frac=0.97, hard_atol=5.0, what="reg(bf16)")


def test_head_vovnet_shapes_vs_oracle():
    """BASELINE config 4 (`detr3d_vovnet_gridmask_det_final_trainval_cbgs`: FPN start_level=0, levels 232x400 ... 29x50,
    4x the res101 texels): the whole head in fp32 parity mode and in bf16 against the oracle on the same inputs."""
    from transcar_b200 import plugin
    Q, B, seed = 900, 1, 7
    sd = synthetic.make_state_dict(seed=seed, num_query=Q)
    feats = synthetic.make_feats(seed, B, "vovnet", smooth=True)
    metas = synthetic.make_img_metas(B, seed=seed)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = O.head_forward(sd, feats, metas)
    for precision, atol, rtol, frac, hard in (("fp32", 5e-5, 1e-5, 0.99, None), ("bf16x3", 1e-3, 1e-2, 0.99, 2.0),
                                              ("bf16", 2e-2, 1e-2, 0.93, 5.0)):
        cfg = synthetic.head_config(num_query=Q)
        cfg["precision"] = precision
        head = plugin.build_head(cfg)
        head.load_state_dict(sd, strict=True)
        head = head.cuda().eval()
        with torch.no_grad():
            got = head([f.cuda() for f in feats], metas)
        torch.cuda.synchronize()
        for k in ("all_cls_scores", "all_bbox_preds"):
            assert_close_tail(got[k].float().cpu().numpy(), want[k].numpy(), atol=atol, rtol=rtol, frac=frac, hard_atol=hard,
                              what=f"vovnet {precision} {k}")
        del head
        torch.cuda.empty_cache()


def test_cross_atten_module_dropin():
    """`Detr3DCrossAtten.forward` with the reference signature vs the oracle's cross_atten (fp32, 1e-5)."""
    from transcar_b200 import plugin
    Q, B = 128, 2
    sd = synthetic.make_state_dict(seed=4, num_query=Q)
    pre = "transformer.decoder.layers.3.attentions.1."
    mod = plugin.ATTENTION.build(dict(type="Detr3DCrossAtten", pc_range=synthetic.PC_RANGE, num_points=1, embed_dims=256))
    mod.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, strict=True)
    mod = mod.cuda().eval()
    feats = synthetic.make_feats(4, B, "tiny", smooth=True)
    metas = synthetic.make_img_metas(B, seed=4)
    g = torch.Generator().manual_seed(8)
    query, pos = torch.randn((Q, B, 256), generator=g), torch.randn((Q, B, 256), generator=g)
    ref = torch.rand((B, Q, 3), generator=g)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = O.cross_atten(sd, pre[:-1], query, pos, feats, ref, metas)
        got = mod(query.cuda(), None, [f.cuda() for f in feats], query_pos=pos.cuda(), reference_points=ref.cuda(),
                  img_metas=metas)
    torch.testing.assert_close(got.cpu(), want, rtol=1e-5, atol=2e-5)
    _, _, mask = O.feature_sampling(feats, ref, metas)
    assert torch.equal(mod.last_mask.bool().cpu(), mask[:, 0, :, :, 0, 0])


def test_cross_atten_num_points_5_and_mmcv_attention_masks():
    """Boundary completeness: (a) `Detr3DCrossAtten(num_points=5)` - the reference default - one sampled point broadcast
    against five weights (T:346-373); (b) the mmcv `MultiheadAttention` wrapper with `attn_mask` / `key_padding_mask`
    against torch's nn.MultiheadAttention."""
    from transcar_b200 import plugin
    Q, B = 96, 2
    gen = torch.Generator().manual_seed(31)
    mod = plugin.ATTENTION.build(dict(type="Detr3DCrossAtten", pc_range=synthetic.PC_RANGE, num_points=5, embed_dims=256))
    with torch.no_grad():
        mod.attention_weights.weight.copy_(torch.randn(mod.attention_weights.weight.shape, generator=gen) * 0.05)
    mod = mod.cuda().eval()
    pre = "x"
    sd = {pre + "." + k: v.detach() for k, v in mod.state_dict().items()}      # oracle on the GPU (ATen's CUDA grid_sample)
    feats = synthetic.make_feats(6, B, "tiny", smooth=True)
    metas = synthetic.make_img_metas(B, seed=6)
    query, pos = torch.randn((Q, B, 256), generator=gen), torch.randn((Q, B, 256), generator=gen)
    ref = torch.rand((B, Q, 3), generator=gen)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = O.cross_atten(sd, pre, query.cuda(), pos.cuda(), [f.cuda() for f in feats], ref.cuda(), metas)
        got = mod(query.cuda(), None, [f.cuda() for f in feats], query_pos=pos.cuda(), reference_points=ref.cuda(), img_metas=metas)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=4e-5)
    # ---- (b)
    att = plugin.ATTENTION.build(dict(type="MultiheadAttention", embed_dims=256, num_heads=8, dropout=0.1)).cuda().eval()
    ref_mha = torch.nn.MultiheadAttention(256, 8).cuda().eval()
    ref_mha.load_state_dict(att.attn.state_dict())
    Lq, Lk = 70, 90
    q, k = torch.randn((Lq, B, 256), generator=gen).cuda(), torch.randn((Lk, B, 256), generator=gen).cuda()
    amask = torch.rand((Lq, Lk), generator=gen).cuda() < 0.3
    amask[:, 0] = False                                           # every row keeps at least one key
    kpm = torch.rand((B, Lk), generator=gen).cuda() < 0.2
    kpm[:, 0] = False
    with torch.no_grad():
        got = att(q, k, k, attn_mask=amask, key_padding_mask=kpm)
        want = q + ref_mha(q, k, k, attn_mask=amask, key_padding_mask=kpm)[0]
    torch.testing.assert_close(got, want, rtol=1e-5, atol=2e-5)


def test_transformer_module_dropin():
    """`Detr3DTransformer.forward(mlvl_feats, query_embed, reg_branches, img_metas=...)` return contract."""
    g, head, sd, feats, metas = build("tiny", "fp32")
    B, Q = int(g["batch"]), int(g["num_query"])
    with torch.no_grad():
        hs, init_ref, inter_refs = head.transformer([f.cuda() for f in feats], head.query_embedding.weight,
                                                    reg_branches=head.reg_branches, img_metas=metas)
    assert hs.shape == (6, Q, B, 256) and init_ref.shape == (B, Q, 3) and inter_refs.shape == (6, B, Q, 3)
    for b in range(B):
        np.testing.assert_allclose(hs[0][:, b].cpu().numpy(), g[f"b{b}.dec0"], rtol=0, atol=1e-5)


def test_decoder_module_dropin():
    """`Detr3DTransformerDecoder.forward(query, key, value, query_pos=, reference_points=, reg_branches=, img_metas=)`
    (T:155-214) from caller-given queries / reference points - not the learned embedding - vs the oracle's layer loop."""
    g, head, sd, feats, metas = build("tiny", "fp32")
    B, Q = 2, 96
    gen = torch.Generator().manual_seed(21)
    query, pos = torch.randn((Q, B, 256), generator=gen), torch.randn((Q, B, 256), generator=gen)
    ref0 = torch.rand((B, Q, 3), generator=gen) * 0.8 + 0.1
    feats = synthetic.make_feats(5, B, "tiny", smooth=True)
    metas = synthetic.make_img_metas(B, seed=5)
    dec = head.transformer.decoder
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        x, ref, want_hs, want_refs = query, ref0, [], []
        for lid in range(6):
            x = O.decoder_layer(sd, f"transformer.decoder.layers.{lid}", x, pos, feats, ref, metas)
            tmp = O.reg_branch(sd, f"reg_branches.{lid}", x.permute(1, 0, 2))
            new = torch.zeros_like(ref)
            new[..., :2] = tmp[..., :2] + O.logit(ref[..., :2])
            new[..., 2:3] = tmp[..., 4:5] + O.logit(ref[..., 2:3])
            ref = new.sigmoid()
            want_hs.append(x)
            want_refs.append(ref)
        hs, refs = dec(query.cuda(), None, [f.cuda() for f in feats], query_pos=pos.cuda(), reference_points=ref0.cuda(),
                       reg_branches=head.reg_branches, img_metas=metas)
        assert hs.shape == (6, Q, B, 256) and refs.shape == (6, B, Q, 3)
        np.testing.assert_allclose(hs[0].cpu().numpy(), want_hs[0].numpy(), rtol=0, atol=2e-5)
        assert_close_tail(hs[5].cpu().numpy(), want_hs[5].numpy(), atol=5e-5, rtol=1e-5, frac=0.99, what="decoder hs5")
        assert_close_tail(refs[5].cpu().numpy(), want_refs[5].numpy(), atol=5e-5, rtol=1e-5, frac=0.99, what="decoder refs5")
        # reg_branches=None: reference points are not refined (T:190)
        hs2, refs2 = dec(query.cuda(), None, [f.cuda() for f in feats], query_pos=pos.cuda(), reference_points=ref0.cuda(),
                         reg_branches=None, img_metas=metas)
        assert torch.equal(refs2[5].cpu(), ref0)
        np.testing.assert_allclose(hs2[0].cpu().numpy(), want_hs[0].numpy(), rtol=0, atol=2e-5)


def test_feature_sampling_free_function():
    """`plugin.feature_sampling` (the reference's free function T:381-422): un-reduced, un-masked samples + bool mask."""
    from transcar_b200 import plugin
    B, Q = 2, 200
    feats = synthetic.make_feats(4, B, "tiny", smooth=True)
    metas = synthetic.make_img_metas(B, seed=4)
    gen = torch.Generator().manual_seed(8)
    ref = torch.rand((B, Q, 3), generator=gen) * 1.2 - 0.1
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want_ref, want_s, want_m = O.feature_sampling([f.cuda() for f in feats], ref.cuda(), metas)
        got_ref, got_s, got_m = plugin.feature_sampling([f.cuda() for f in feats], ref.cuda(), synthetic.PC_RANGE, metas)
    assert got_s.shape == want_s.shape == (B, 256, Q, 6, 1, 4) and got_m.shape == want_m.shape and got_m.dtype == torch.bool
    assert torch.equal(got_m, want_m) and torch.equal(got_ref, want_ref)
    assert (~want_m).any() and (want_s[(~want_m).expand(B, 256, Q, 6, 1, 1).expand_as(want_s)] != 0).any(), \
        "the case must contain cameras that fail the validity test but still sample non-zero texels"
    torch.testing.assert_close(got_s, torch.nan_to_num(want_s), rtol=0, atol=2e-5)


def test_get_bboxes_vs_reference_decode():
    g, head, sd, feats, metas = build("tiny", "fp32")
    preds = dict(all_cls_scores=torch.from_numpy(g["all_cls_scores"]).cuda(),
                 all_bbox_preds=torch.from_numpy(g["all_bbox_preds"]).cuda())
    res = head.get_bboxes(preds, metas)
    for b, (bboxes, scores, labels) in enumerate(res):
        want = g[f"b{b}.decode.bboxes"].copy()
        want[:, 2] = want[:, 2] - want[:, 5] * 0.5
        np.testing.assert_allclose(bboxes.cpu().numpy(), want, rtol=1e-6, atol=1e-5)
        np.testing.assert_allclose(scores.cpu().numpy(), g[f"b{b}.decode.scores"], rtol=0, atol=1e-6)
        assert np.array_equal(labels.cpu().numpy(), g[f"b{b}.decode.labels"])


def test_batch8_equals_per_sample():
    """H7: batched execution == looping samples (bit-identical: no cross-sample interaction)."""
    from transcar_b200 import plugin
    Q = 128
    sd = synthetic.make_state_dict(seed=2, num_query=Q)
    cfg = synthetic.head_config(num_query=Q)
    cfg["precision"] = "fp32"
    head = plugin.build_head(cfg)
    head.load_state_dict(sd)
    head = head.cuda().eval()
    B = 8
    feats = [f.cuda() for f in synthetic.make_feats(2, B, "tiny")]
    metas = synthetic.make_img_metas(B, seed=2)
    metas = [dict(m, lidar2img=metas[0]["lidar2img"]) for m in metas] if False else metas
    with torch.no_grad():
        full = head(feats, metas)
        for b in (0, 3, 7):
            one = head([f[b:b + 1] for f in feats], [dict(metas[b], img_shape=metas[0]["img_shape"])])
            assert torch.equal(one["all_cls_scores"][:, 0], full["all_cls_scores"][:, b])
            assert torch.equal(one["all_bbox_preds"][:, 0], full["all_bbox_preds"][:, b])
