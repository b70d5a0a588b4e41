"""Seeded synthetic workloads for the fusion-decoder hot path.

Everything here is deterministic in (seed, shape) through *CPU* generators, so the
build container, the GPU box, the oracle, the tests and ``bench.py`` all see the same
numbers.  Nothing here touches ``/root/reference`` or ``oracle/``.

* :func:`head_config` - the ``pts_bbox_head`` dict of the reference configs
  (``projects/configs/detr3d/detr3d_res101_gridmask.py:51-102``), same keys.
* :func:`make_state_dict` - random weights under the *reference's* state-dict names
  (SURVEY.md section 8b); loading it ``strict=True`` is the checkpoint-compat test.
* :func:`make_feats`, :func:`make_img_metas` - FPN features, real 6-camera calibration
  (fixture ``tests/golden/nus_calib.npz``), synthetic radar sweeps + 36-d tokens.
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np
import torch

from .radar_tokens import RADAR_CHANNELS, build_radar_tokens

PC_RANGE = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
IMG_SHAPE = (928, 1600, 3)          # padded shape (quirk Q2: H is the padded 928)

LEVEL_SHAPES = {
    # FPN outputs for the padded 928x1600 input
    "res101": [(116, 200), (58, 100), (29, 50), (15, 25)],      # start_level=1
    "vovnet": [(232, 400), (116, 200), (58, 100), (29, 50)],    # start_level=0
    "tiny": [(16, 28), (8, 14), (4, 7), (2, 4)],                # CPU-sized test case
}

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CALIB_PATH = os.path.join(_REPO_ROOT, "tests", "golden", "nus_calib.npz")


def head_config(num_query=900):
    """``pts_bbox_head`` of ``detr3d_res101_gridmask.py:51-102`` (train/test cfg omitted)."""
    return dict(
        type="Detr3DHead",
        num_query=num_query,
        num_classes=10,
        in_channels=256,
        sync_cls_avg_factor=True,
        with_box_refine=True,
        as_two_stage=False,
        transformer=dict(
            type="Detr3DTransformer",
            decoder=dict(
                type="Detr3DTransformerDecoder",
                num_layers=6,
                return_intermediate=True,
                transformerlayers=dict(
                    type="DetrTransformerDecoderLayer",
                    attn_cfgs=[
                        dict(type="MultiheadAttention", embed_dims=256, num_heads=8, dropout=0.1),
                        dict(type="Detr3DCrossAtten", pc_range=PC_RANGE, num_points=1, embed_dims=256),
                    ],
                    feedforward_channels=512,
                    ffn_dropout=0.1,
                    operation_order=("self_attn", "norm", "cross_attn", "norm", "ffn", "norm")))),
        bbox_coder=dict(
            type="NMSFreeCoder",
            post_center_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0],
            pc_range=PC_RANGE,
            max_num=300,
            voxel_size=[0.2, 0.2, 8],
            num_classes=10),
        positional_encoding=dict(type="SinePositionalEncoding", num_feats=128, normalize=True, offset=-0.5),
        loss_cls=dict(type="FocalLoss", use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0),
        loss_bbox=dict(type="L1Loss", loss_weight=0.25),
        loss_iou=dict(type="GIoULoss", loss_weight=0.0),
    )


# --------------------------------------------------------------------------- weights
def _linear(names, prefix, n_out, n_in):
    names.append((prefix + ".weight", (n_out, n_in), "w"))
    names.append((prefix + ".bias", (n_out,), "b"))


def _norm(names, prefix, n):
    names.append((prefix + ".weight", (n,), "g"))
    names.append((prefix + ".bias", (n,), "b"))


def _mha(names, prefix, e):
    names.append((prefix + ".in_proj_weight", (3 * e, e), "w"))
    names.append((prefix + ".in_proj_bias", (3 * e,), "b"))
    _linear(names, prefix + ".out_proj", e, e)


def _cls_branch(names, prefix, e, n_cls):
    _linear(names, prefix + ".0", e, e); _norm(names, prefix + ".1", e)
    _linear(names, prefix + ".3", e, e); _norm(names, prefix + ".4", e)
    _linear(names, prefix + ".6", n_cls, e)


def _reg_branch(names, prefix, e, code):
    _linear(names, prefix + ".0", e, e)
    _linear(names, prefix + ".2", e, e)
    _linear(names, prefix + ".4", code, e)


def state_dict_spec(num_query=900, embed=256, ffn=512, num_layers=6, num_cls=10, code=10,
                    cams=6, levels=4, points=1):
    """(name, shape, kind) for every tensor of the reference ``Detr3DHead`` state dict.

    kind: ``w`` weight matrix, ``b`` bias, ``g`` LayerNorm gain, ``e`` embedding, ``c`` code weights.
    Names per SURVEY.md section 8b (``detr3d_head.py:70-238``, mmcv layer wrapper keys).
    """
    names = []
    names.append(("code_weights", (code,), "c"))
    for i in range(num_layers):
        p = f"transformer.decoder.layers.{i}."
        _mha(names, p + "attentions.0.attn", embed)
        _linear(names, p + "attentions.1.attention_weights", cams * levels * points, embed)
        _linear(names, p + "attentions.1.output_proj", embed, embed)
        _linear(names, p + "attentions.1.position_encoder.0", embed, 3)
        _norm(names, p + "attentions.1.position_encoder.1", embed)
        _linear(names, p + "attentions.1.position_encoder.3", embed, embed)
        _norm(names, p + "attentions.1.position_encoder.4", embed)
        _linear(names, p + "ffns.0.layers.0.0", ffn, embed)
        _linear(names, p + "ffns.0.layers.1", embed, ffn)
        for k in range(3):
            _norm(names, p + f"norms.{k}", embed)
    _linear(names, "transformer.reference_points", 3, embed)
    for i in range(num_layers):
        _cls_branch(names, f"cls_branches.{i}", embed, num_cls)
    for i in range(num_layers):
        _reg_branch(names, f"reg_branches.{i}", embed, code)
    names.append(("query_embedding.weight", (num_query, 2 * embed), "e"))
    for s in ("", "2", "3"):
        _cls_branch(names, "final_cls" + s, embed, num_cls)
        _reg_branch(names, "final_reg" + s, embed, code)
    for s in ("", "_2", "_3"):
        _mha(names, "rf_multihead_attn" + s.replace("_", ""), embed)
        _linear(names, "rf_linear1" + s, ffn, embed)
        _linear(names, "rf_linear2" + s, embed, ffn)
        for k in (1, 2, 3):
            _norm(names, f"rf_norm{k}" + s, embed)
    _linear(names, "radar_position_encoder.0", embed, 3)
    _norm(names, "radar_position_encoder.1", embed)
    _linear(names, "radar_position_encoder.3", embed, embed)
    _norm(names, "radar_position_encoder.4", embed)
    _linear(names, "radar_feat_encoder.0", 64, 36)
    _linear(names, "radar_feat_encoder.2", 128, 64)
    _linear(names, "radar_feat_encoder.4", embed, 128)
    for s in ("2", "3"):                                  # never used in forward, but in checkpoints
        _linear(names, "attention_weights" + s, cams * levels, embed)
        _linear(names, "output_proj" + s, embed, embed)
    return names


def make_state_dict(seed=0, num_query=900, dtype=torch.float32):
    """Random-init weights under reference names.  Xavier-uniform matrices, small biases,
    LayerNorm gains around 1; ``attention_weights`` ~ N(0, 0.05) so the sigmoid camera/level
    weights are not the constant 0.5 the reference's zero init gives (quirk Q8)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * (seed + 1))
    sd = OrderedDict()
    for name, shape, kind in state_dict_spec(num_query=num_query):
        if kind == "w":
            fan_out, fan_in = shape
            a = (6.0 / (fan_in + fan_out)) ** 0.5
            t = (torch.rand(shape, generator=g) * 2 - 1) * a
            if name.endswith("attention_weights.weight"):
                t = torch.randn(shape, generator=g) * 0.05
            if name.endswith("position_encoder.0.weight") and name.startswith("radar"):
                t = t * 0.05     # radar xyz are metres (|x| up to 51 and the 500 pad): keep pre-LN values finite and tame
            if name == "radar_feat_encoder.0.weight":
                t = t * 0.2
            # regression heads: small last layer, like a trained box head emitting modest deltas.  With
            # unit-gain deltas the 6-layer "refine reference point -> resample" loop is chaotic (measured:
            # x30-200 error growth per layer), which would make any end-to-end comparison meaningless.
            if name.endswith(".4.weight") and ("reg" in name):
                t = t * 0.05
        elif kind == "b":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
            if name.endswith(".4.bias") and ("reg" in name):
                # box code (cx,cy,w,l,cz,h,sin,cos,vx,vy): length exp(1.0)=2.7 m -> attention radius 1.36 m
                # (inside the [1,2] clamp, above the [0.5,1] one); heading terms move the front/rear circles
                t[3] += 1.0
                t[6] += 0.6
                t[7] -= 0.5
        elif kind == "g":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "e":
            t = torch.randn(shape, generator=g)
        elif kind == "c":
            t = torch.tensor([1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.2, 0.2])
        else:
            raise AssertionError(kind)
        sd[name] = t.to(dtype).contiguous()
    return sd


def state_dict_checksum(sd):
    """Order-sensitive float64 checksum, used by the golden fixtures to detect RNG drift."""
    acc = 0.0
    for i, (k, v) in enumerate(sd.items()):
        acc += float(v.double().sum()) * (1.0 + 1e-3 * (i % 97)) + float(v.double().abs().sum()) * 1e-2
    return acc


# --------------------------------------------------------------------------- inputs
def load_calibration():
    """``[2,6,4,4]`` float64 ``lidar2img`` of the two nuScenes infos in the reference's own test
    fixture (``mmdetection3d/tests/data/nuscenes/nus_info.pkl``), produced by
    ``oracle/gen_golden.py`` with the formula at ``mmdet3d/datasets/nuscenes_dataset.py:223-236``."""
    if not os.path.exists(_CALIB_PATH):
        raise FileNotFoundError(f"calibration fixture missing: {_CALIB_PATH} (run oracle/gen_golden.py)")
    return np.load(_CALIB_PATH)["lidar2img"]


def make_feats(seed, batch, config="res101", channels=256, cams=6, device="cpu", dtype=torch.float32,
               channels_last=False, smooth=True):
    """Four FPN levels ``[B,cams,C,H_l,W_l]``, unit scale.

    ``smooth=True`` (end-to-end cases): band-limited features - white noise on a 128-pixel image grid
    (9x14 knots) bilinearly upsampled to every level plus 5 % white noise - because real FPN features
    are smooth at the stride scale while white-noise maps make the iterative reference-point refinement
    chaotic.  ``smooth=False`` (single-stage sampling tests, bandwidth bench): plain N(0,1) texels.
    With ``channels_last=True`` the *logical* shape is unchanged but memory is ``[B,cams,H,W,C]``
    (what a channels-last backbone hands over; zero-copy input layout of the sampling kernel)."""
    shapes = LEVEL_SHAPES[config] if isinstance(config, str) else config
    g = torch.Generator(device="cpu")
    g.manual_seed(7919 * (seed + 1) + 13)
    feats = []
    for (h, w) in shapes:
        if smooth:
            knots = torch.randn((batch * cams, channels, 9, 14), generator=g, dtype=torch.float32)
            t = torch.nn.functional.interpolate(knots, size=(h, w), mode="bilinear", align_corners=True)
            t = (1.25 * t + 0.05 * torch.randn(t.shape, generator=g)).view(batch, cams, channels, h, w)
        else:
            t = torch.randn((batch, cams, channels, h, w), generator=g, dtype=torch.float32)
        t = t.to(device=device, dtype=dtype)
        if channels_last:
            t = t.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
        feats.append(t)
    return feats


_RADAR_YAW = {"LIDAR_TOP": -1.57, "RADAR_FRONT": 0.0, "RADAR_FRONT_LEFT": 1.54,
              "RADAR_FRONT_RIGHT": -1.59, "RADAR_BACK_LEFT": 3.04, "RADAR_BACK_RIGHT": -3.07}


def _yaw_matrix(a):
    return np.array([[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]])


def make_radar_sweeps(sample_seed, n_per_channel=400):
    """Synthetic devkit-shaped radar returns for one sample: per channel ``points [18,n]``
    (field order at ``detr3d_head.py:498``), ``lags [1,n]`` seconds, sensor rotation."""
    sweeps = {"LIDAR_TOP": dict(rotation=_yaw_matrix(_RADAR_YAW["LIDAR_TOP"]))}
    for ci, chan in enumerate(RADAR_CHANNELS):
        rs = np.random.RandomState(100003 * (sample_seed + 1) + 17 * ci)
        n = n_per_channel
        p = np.zeros((18, n))
        p[0] = rs.uniform(-60, 60, n)
        p[1] = rs.uniform(-60, 60, n)
        p[2] = rs.uniform(-1, 1, n)
        p[3] = rs.randint(0, 8, n)
        p[4] = rs.randint(0, 100, n)
        p[5] = rs.uniform(-5, 30, n)
        p[6:10] = rs.randn(4, n) * 3
        p[10] = 1
        p[11] = rs.randint(0, 5, n)
        p[12:14] = rs.randint(0, 20, (2, n))
        p[14] = rs.randint(0, 18, n)
        p[15] = rs.randint(0, 8, n)
        p[16:18] = rs.randint(0, 20, (2, n))
        lags = rs.choice([0.0, 0.075, 0.15, 0.225, 0.3], n)[None, :]
        sweeps[chan] = dict(points=p, lags=lags, rotation=_yaw_matrix(_RADAR_YAW[chan]))
    return sweeps


def make_img_metas(batch, seed=0, n_per_channel=400, with_sweeps=False):
    """``img_metas`` list for ``Detr3DHead.forward``: ``lidar2img`` (6 float64 4x4), ``img_shape``,
    ``sample_idx`` and the I/O-free extension key ``radar_tokens`` (``[n,36]`` float32)."""
    calib = load_calibration()
    metas = []
    for b in range(batch):
        sample_seed = seed * 1000 + b
        sweeps = make_radar_sweeps(sample_seed, n_per_channel)
        meta = dict(
            lidar2img=[calib[b % calib.shape[0], c].copy() for c in range(calib.shape[1])],
            img_shape=[IMG_SHAPE] * calib.shape[1],
            sample_idx=f"synthetic-{sample_seed}",
            radar_tokens=build_radar_tokens(sweeps),
        )
        if with_sweeps:
            meta["radar_sweeps"] = sweeps
        metas.append(meta)
    return metas
