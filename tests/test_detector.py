"""N2: backbone + FPN hand-off into the fusion decoder - VoVNet-99 / FPN level shapes for both TransCAR
configurations (CPU) and the zero-copy channels-last bf16 hand-off + full-model forward (``-m gpu``)."""
import warnings

import pytest
import torch

from transcar_b200 import detector, ops, synthetic


def test_vovnet_fpn_level_shapes_cpu():
    """Strides 4/8/16/32 out of VoVNet; FPN start_level=1 + one extra stride-2 level = the res101 config's level shapes
    (here for a 128 x 224 image: the 'tiny' shapes), start_level=0 = the VoVNet config's."""
    torch.manual_seed(0)
    net = detector.VoVNet("V-39-eSE").eval()                      # same code path as V-99, 2.5x fewer modules
    x = torch.randn(2, 3, 128, 224)
    with torch.no_grad():
        feats = net(x)
        assert [tuple(f.shape[1:]) for f in feats] == [(256, 32, 56), (512, 16, 28), (768, 8, 14), (1024, 4, 7)]
        a = detector.FPN(net.out_channels, 256, num_outs=4, start_level=1).eval()(feats)
        b = detector.FPN(net.out_channels, 256, num_outs=4, start_level=0).eval()(feats)
    assert [tuple(f.shape[2:]) for f in a] == synthetic.LEVEL_SHAPES["tiny"]
    assert [tuple(f.shape[2:]) for f in b] == [(32, 56), (16, 28), (8, 14), (4, 7)]
    assert all(f.shape[1] == 256 for f in a + b)
    spec = detector.SPECS["V-99-eSE"]
    assert spec["blocks"] == (1, 3, 9, 3) and spec["layers"] == 5          # cfg ...trainval_cbgs.py:33-38
    n99 = sum(p.numel() for p in detector.VoVNet("V-99-eSE").parameters())
    assert 60e6 < n99 < 80e6            # 69.5 M parameters in the backbone


def test_extract_img_feat_layout_cpu():
    det = detector.Detr3D(detector.VoVNet("V-39-eSE"), detector.FPN([256, 512, 768, 1024], 256, 4, 1), None,
                          handoff_dtype=torch.float32).eval()
    metas = [dict(), dict()]
    with torch.no_grad():
        feats = det.extract_img_feat(torch.randn(2, 6, 3, 64, 96), metas)
    assert metas[0]["input_shape"] == (64, 96)
    for f in feats:
        assert f.shape[:3] == (2, 6, 256) and ops.is_channels_last_5d(f)


@pytest.mark.gpu
def test_full_model_zero_copy_handoff_and_parity():
    """Images -> VoVNet-99 + FPN (cuDNN, channels-last bf16 autocast) -> fusion decoder.  The decoder reads the FPN
    outputs in place (same storage, channels-last bf16), and its result matches the oracle run on those very maps."""
    from oracle import fusion_decoder as O
    Q, B = 128, 2
    torch.manual_seed(0)
    cfg = synthetic.head_config(num_query=Q)
    det = detector.build_detector(cfg, "V-99-eSE", start_level=1)
    sd = synthetic.make_state_dict(seed=3, num_query=Q)
    det.pts_bbox_head.load_state_dict(sd, strict=True)
    img = torch.randn(B, 6, 3, 128, 224, device="cuda")
    metas = synthetic.make_img_metas(B, seed=3)
    with torch.no_grad():
        feats = det.extract_img_feat(img, metas)
        assert [tuple(f.shape) for f in feats] == [(B, 6, 256, h, w) for h, w in synthetic.LEVEL_SHAPES["tiny"]]
        for f in feats:
            assert f.dtype == torch.bfloat16 and ops.is_channels_last_5d(f) and torch.isfinite(f.float()).all()
        eng = det.pts_bbox_head.engine()
        prepared = eng.prepare_inputs(feats, metas)
        assert all(p.data_ptr() == f.data_ptr() for p, f in zip(prepared[0], feats)), "hand-off must be zero-copy"
        got = det.pts_bbox_head(feats, metas)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = O.head_forward({k: v.cuda() for k, v in sd.items()}, [f.float() for f in feats], metas)
        res = det.simple_test(metas, img)
    from parity import assert_close_tail
    for k in ("all_cls_scores", "all_bbox_preds"):
        assert_close_tail(got[k].cpu().numpy(), want[k].cpu().numpy(), atol=1e-3, rtol=1e-2, frac=0.99, hard_atol=2.0, what=k)
    assert len(res) == B and res[0]["pts_bbox"]["boxes_3d"].shape[1] == 9
