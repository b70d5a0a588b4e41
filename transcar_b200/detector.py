"""Image side of the full model: backbone + FPN producing the four feature levels the
fusion decoder samples, handed over ZERO-COPY in the layout and dtype the sampling kernel reads.

Reference call path: ``projects/mmdet3d_plugin/models/detectors/detr3d.py:39-66`` (``Detr3D.extract_img_feat``: the six
camera images of a sample are flattened to ``[B*N, 3, H, W]``, GridMask - identity in eval - backbone, neck, and every
level is viewed back as ``[B, N, C, H_l, W_l]``) and ``:169-176`` (``simple_test_pts``: head forward + ``get_bboxes``).

The convolutions are NOT rewritten here (out of scope - they stay cuDNN through PyTorch); what this
module owns is the hand-off: the backbone and the neck run under ``torch.channels_last`` + bf16 autocast, so every FPN
output ``[B*N, 256, H, W]`` is physically ``[B*N, H, W, 256]`` bf16, and its ``[B, N, 256, H, W]`` view is exactly the
channels-last layout ``tc_sample_fwd`` gathers from (one texel = 512 contiguous bytes): no NCHW->NHWC pass, no cast, no
copy between the neck and the decoder.

* :class:`VoVNet` - VoVNet-V2 (One-Shot-Aggregation modules with effective Squeeze-Excitation), spec ``V-99-eSE``, the
  backbone of ``projects/configs/detr3d/detr3d_vovnet_gridmask_det_final_trainval_cbgs.py:32-38`` (reference module:
  ``models/backbones/vovnet.py:269-374``), written from the published architecture with plain ``torch.nn`` modules.
* :class:`FPN` - mmdet's FPN for the two configurations TransCAR uses (``start_level`` 0 or 1, ``add_extra_convs=
  'on_output'``, ``relu_before_extra_convs=True``, 4 outputs): cfg ``detr3d_res101_gridmask.py:43-50`` / ``...cbgs.py:39-46``.
* :class:`Detr3D` - the detector shell: ``extract_img_feat`` / ``simple_test`` with the reference's signatures.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

# published VoVNet-V2 specifications (stem widths, per-stage 3x3 width, per-stage output width, 3x3 layers per OSA
# module, OSA modules per stage)
SPECS = {
    "V-39-eSE": dict(stem=(64, 64, 128), conv_ch=(128, 160, 192, 224), out_ch=(256, 512, 768, 1024), layers=5, blocks=(1, 1, 2, 2)),
    "V-57-eSE": dict(stem=(64, 64, 128), conv_ch=(128, 160, 192, 224), out_ch=(256, 512, 768, 1024), layers=5, blocks=(1, 1, 4, 3)),
    "V-99-eSE": dict(stem=(64, 64, 128), conv_ch=(128, 160, 192, 224), out_ch=(256, 512, 768, 1024), layers=5, blocks=(1, 3, 9, 3)),
}


def _conv_bn_relu(cin, cout, k, stride=1):
    return nn.Sequential(nn.Conv2d(cin, cout, k, stride, k // 2, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class _ESE(nn.Module):
    """effective Squeeze-Excitation: x * hard_sigmoid(fc(global_avg_pool(x))), one 1x1 conv, no reduction."""

    def __init__(self, ch):
        super().__init__()
        self.fc = nn.Conv2d(ch, ch, 1)

    def forward(self, x):
        return x * F.hardsigmoid(self.fc(x.mean((2, 3), keepdim=True)))


class _OSA(nn.Module):
    """One-Shot Aggregation: `layers` chained 3x3 convs, ALL intermediate maps concatenated once, 1x1 fusion, eSE,
    identity shortcut from the second module of a stage on."""

    def __init__(self, cin, conv_ch, cout, layers, identity):
        super().__init__()
        self.identity = identity
        self.layers = nn.ModuleList(_conv_bn_relu(cin if i == 0 else conv_ch, conv_ch, 3) for i in range(layers))
        self.concat = _conv_bn_relu(cin + layers * conv_ch, cout, 1)
        self.ese = _ESE(cout)

    def forward(self, x):
        outs = [x]
        y = x
        for layer in self.layers:
            y = layer(y)
            outs.append(y)
        y = self.ese(self.concat(torch.cat(outs, 1)))
        return y + x if self.identity else y


class VoVNet(nn.Module):
    def __init__(self, spec_name="V-99-eSE", input_ch=3, out_features=("stage2", "stage3", "stage4", "stage5"),
                 norm_eval=True, frozen_stages=-1, **kwargs):
        super().__init__()
        s = SPECS[spec_name]
        self.out_features, self.norm_eval, self.frozen_stages = tuple(out_features), norm_eval, frozen_stages
        st = s["stem"]
        self.stem = nn.Sequential(_conv_bn_relu(input_ch, st[0], 3, 2), _conv_bn_relu(st[0], st[1], 3, 1),
                                  _conv_bn_relu(st[1], st[2], 3, 2))                       # stride 4
        self.stages = nn.ModuleList()
        cin = st[2]
        for i, (cc, co, nb) in enumerate(zip(s["conv_ch"], s["out_ch"], s["blocks"])):
            mods = [] if i == 0 else [nn.MaxPool2d(3, 2, ceil_mode=True)]                  # stages 3-5 halve the resolution
            for b in range(nb):
                mods.append(_OSA(cin if b == 0 else co, cc, co, s["layers"], identity=b > 0))
            self.stages.append(nn.Sequential(*mods))
            cin = co
        self.out_channels = list(s["out_ch"])

    def forward(self, x):
        x = self.stem(x)
        outs = []
        for i, stage in enumerate(self.stages):
            x = stage(x)
            if f"stage{i + 2}" in self.out_features:
                outs.append(x)
        return outs

    def train(self, mode=True):
        super().train(mode)
        if mode and self.norm_eval:                 # reference: norm_eval=True keeps BatchNorm statistics frozen
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    m.eval()
        return self


class FPN(nn.Module):
    """mmdet FPN restricted to what the TransCAR configs use.  Outputs ``num_outs`` maps of ``out_channels`` channels."""

    def __init__(self, in_channels, out_channels=256, num_outs=4, start_level=0, add_extra_convs="on_output",
                 relu_before_extra_convs=True, **kwargs):
        super().__init__()
        if add_extra_convs not in ("on_output", False, None):
            raise NotImplementedError("transcar_b200.FPN: add_extra_convs must be 'on_output' (the TransCAR configs)")
        self.start_level, self.num_outs, self.relu_before_extra_convs = start_level, num_outs, relu_before_extra_convs
        used = len(in_channels) - start_level
        self.lateral_convs = nn.ModuleList(nn.Conv2d(c, out_channels, 1) for c in in_channels[start_level:])
        self.fpn_convs = nn.ModuleList(nn.Conv2d(out_channels, out_channels, 3, padding=1) for _ in range(used))
        for _ in range(num_outs - used):            # extra levels: stride-2 3x3 convs on the previous OUTPUT
            self.fpn_convs.append(nn.Conv2d(out_channels, out_channels, 3, stride=2, padding=1))
        self.used = used

    def forward(self, feats):
        lat = [conv(feats[i + self.start_level]) for i, conv in enumerate(self.lateral_convs)]
        for i in range(self.used - 1, 0, -1):       # top-down pathway, nearest-neighbour upsampling to the finer level's size
            lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode="nearest")
        outs = [self.fpn_convs[i](lat[i]) for i in range(self.used)]
        for i in range(self.used, self.num_outs):
            src = outs[-1]
            if self.relu_before_extra_convs and i > self.used:
                src = F.relu(src)
            outs.append(self.fpn_convs[i](src))
        return outs


class Detr3D(nn.Module):
    """Detector shell with the reference's ``extract_img_feat(img, img_metas)`` / ``simple_test(img_metas, img)``
    (``detr3d.py:39-66, 169-189``).  ``pts_bbox_head`` is a :class:`transcar_b200.plugin.Detr3DHead`."""

    def __init__(self, img_backbone, img_neck, pts_bbox_head, use_grid_mask=True, handoff_dtype=torch.bfloat16):
        super().__init__()
        self.img_backbone, self.img_neck, self.pts_bbox_head = img_backbone, img_neck, pts_bbox_head
        self.use_grid_mask = use_grid_mask          # GridMask (detr3d.py:37, 55-56) only acts in training mode
        self.handoff_dtype = handoff_dtype

    def extract_img_feat(self, img, img_metas=None):
        """img ``[B, N, 3, H, W]`` (or ``[B*N, 3, H, W]`` with B = 1) -> 4 x logical ``[B, N, 256, H_l, W_l]`` stored
        channels-last in ``handoff_dtype`` - the decoder's zero-copy input."""
        if img.dim() == 5:
            B, N = img.shape[:2]
            img = img.reshape(B * N, *img.shape[2:])
        else:
            B, N = 1, img.shape[0]
        if img_metas is not None:
            for m in img_metas:
                m.update(input_shape=tuple(img.shape[-2:]))
        if self.training and self.use_grid_mask:
            raise NotImplementedError("transcar_b200.Detr3D: GridMask augmentation (training data pipeline) is out of scope")
        img = img.contiguous(memory_format=torch.channels_last)
        with torch.autocast("cuda", dtype=self.handoff_dtype, enabled=self.handoff_dtype != torch.float32):
            feats = self.img_neck(self.img_backbone(img))
        out = []
        for f in feats:
            if f.dtype != self.handoff_dtype or not f.is_contiguous(memory_format=torch.channels_last):
                f = f.to(self.handoff_dtype).contiguous(memory_format=torch.channels_last)     # not taken by the convs above
            out.append(f.view(B, N, *f.shape[1:]))
        return out

    def forward(self, img, img_metas):
        return self.pts_bbox_head(self.extract_img_feat(img, img_metas), img_metas)

    @torch.no_grad()
    def simple_test(self, img_metas, img=None, rescale=False):
        outs = self.forward(img, img_metas)
        return [dict(pts_bbox=dict(boxes_3d=b, scores_3d=s, labels_3d=l))
                for b, s, l in self.pts_bbox_head.get_bboxes(outs, img_metas, rescale=rescale)]


def build_detector(head_cfg, spec_name="V-99-eSE", start_level=1, device="cuda"):
    """VoVNet + FPN + TransCAR fusion head.  ``start_level=1`` reproduces the level shapes of ``detr3d_res101_gridmask``
    (strides 8 / 16 / 32 / 64), ``start_level=0`` those of the VoVNet config (strides 4 / 8 / 16 / 32)."""
    from . import plugin
    backbone = VoVNet(spec_name)
    neck = FPN(backbone.out_channels, 256, num_outs=4, start_level=start_level)
    head = plugin.build_head(head_cfg)
    det = Detr3D(backbone, neck, head)
    return det.to(device=device, memory_format=torch.channels_last).eval()
