"""ORACLE - test infrastructure only.  Import harness for the *unmodified* reference files.

Works only where ``/root/reference`` exists (the build container); nothing on the GPU box uses it.
mmcv / mmdet / nuscenes-devkit / pyquaternion are not installed here, so this module
registers minimal stand-ins in ``sys.modules`` and then executes four reference files verbatim from
where they lie (no copy is made):

    projects/mmdet3d_plugin/core/bbox/util.py
    projects/mmdet3d_plugin/core/bbox/coders/nms_free_coder.py
    projects/mmdet3d_plugin/models/utils/detr3d_transformer.py
    projects/mmdet3d_plugin/models/dense_heads/detr3d_head.py

The mmcv 1.3.8-1.4.0 layer wrapper (``BaseTransformerLayer`` / ``MultiheadAttention`` / ``FFN`` /
``TransformerLayerSequence``) and mmdet's ``DETRHead.__init__`` are restated from their published
behaviour; they only wire ``torch.nn`` modules together - every arithmetic op is
PyTorch's own.  The devkit stand-ins serve the synthetic sweeps registered in :data:`SWEEPS`.
"""
from __future__ import annotations

import copy
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("TRANSCAR_REFERENCE", "/root/reference")
PLUGIN = os.path.join(REFERENCE_ROOT, "projects", "mmdet3d_plugin")

SWEEPS = {}      # sample_idx -> transcar_b200.synthetic.make_radar_sweeps(...) dict


def available():
    return os.path.isfile(os.path.join(PLUGIN, "models", "dense_heads", "detr3d_head.py"))


# ----------------------------------------------------------------------------- registries
class _Registry:
    def __init__(self, name):
        self.name = name
        self.classes = {}

    def register_module(self, *args, **kwargs):
        def wrap(cls):
            self.classes[cls.__name__] = cls
            return cls
        return wrap

    def build(self, cfg, **extra):
        cfg = dict(cfg)
        return self.classes[cfg.pop("type")](**cfg, **extra)


ATTENTION = _Registry("attention")
LAYER_SEQUENCE = _Registry("transformer layer sequence")
TRANSFORMER = _Registry("transformer")
TRANSFORMER_LAYER = _Registry("transformer layer")
HEADS = _Registry("heads")
BBOX_CODERS = _Registry("bbox coders")


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg


# ----------------------------------------------------------------------------- mmcv bricks
class MmcvMultiheadAttention(BaseModule):
    """identity + dropout(nn.MultiheadAttention(q + q_pos, k + k_pos, v))."""

    def __init__(self, embed_dims, num_heads, attn_drop=0.0, proj_drop=0.0, init_cfg=None,
                 batch_first=False, **kwargs):
        super().__init__(init_cfg)
        out_drop = 0.0
        if "dropout" in kwargs:                      # deprecated kwarg: feeds both dropouts
            attn_drop = out_drop = kwargs.pop("dropout")
        self.embed_dims = embed_dims
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
        self.proj_drop = nn.Dropout(proj_drop)
        self.dropout_layer = nn.Dropout(out_drop)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None,
                attn_mask=None, key_padding_mask=None, **kwargs):
        key = query if key is None else key
        value = key if value is None else value
        identity = query if identity is None else identity
        if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
            key_pos = query_pos
        if query_pos is not None:
            query = query + query_pos
        if key_pos is not None:
            key = key + key_pos
        out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask,
                        key_padding_mask=key_padding_mask)[0]
        return identity + self.dropout_layer(self.proj_drop(out))


class MmcvFFN(BaseModule):
    def __init__(self, embed_dims, feedforward_channels, ffn_drop):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True),
                          nn.Dropout(ffn_drop)),
            nn.Linear(feedforward_channels, embed_dims),
            nn.Dropout(ffn_drop))

    def forward(self, x, identity=None):
        return (x if identity is None else identity) + self.layers(x)


class MmcvTransformerLayer(BaseModule):
    """Post-norm ``BaseTransformerLayer``; kwargs (reference_points, img_metas) reach every attention."""

    def __init__(self, attn_cfgs, feedforward_channels, ffn_dropout, operation_order, **kwargs):
        super().__init__()
        self.operation_order = tuple(operation_order)
        self.pre_norm = self.operation_order[0] == "norm"
        assert not self.pre_norm
        self.attentions = nn.ModuleList(ATTENTION.build(c) for c in attn_cfgs)
        self.embed_dims = self.attentions[0].embed_dims
        self.ffns = nn.ModuleList(MmcvFFN(self.embed_dims, feedforward_channels, ffn_dropout)
                                  for _ in range(self.operation_order.count("ffn")))
        self.norms = nn.ModuleList(nn.LayerNorm(self.embed_dims)
                                   for _ in range(self.operation_order.count("norm")))

    def forward(self, query, key=None, value=None, query_pos=None, key_pos=None, attn_masks=None,
                query_key_padding_mask=None, key_padding_mask=None, **kwargs):
        a = f = n = 0
        for op in self.operation_order:
            if op == "self_attn":
                query = self.attentions[a](query, query, query, None, query_pos=query_pos,
                                           key_pos=query_pos, attn_mask=None,
                                           key_padding_mask=query_key_padding_mask, **kwargs)
                a += 1
            elif op == "cross_attn":
                query = self.attentions[a](query, key, value, None, query_pos=query_pos,
                                           key_pos=key_pos, attn_mask=None,
                                           key_padding_mask=key_padding_mask, **kwargs)
                a += 1
            elif op == "norm":
                query = self.norms[n](query)
                n += 1
            elif op == "ffn":
                query = self.ffns[f](query, None)
                f += 1
        return query


class MmcvLayerSequence(BaseModule):
    def __init__(self, transformerlayers=None, num_layers=None, init_cfg=None):
        super().__init__(init_cfg)
        self.num_layers = num_layers
        self.layers = nn.ModuleList(TRANSFORMER_LAYER.build(copy.deepcopy(transformerlayers))
                                    for _ in range(num_layers))
        self.embed_dims = self.layers[0].embed_dims
        self.pre_norm = self.layers[0].pre_norm


class MmdetDETRHead(BaseModule):
    """The ctor side effects of mmdet's ``DETRHead`` that ``Detr3DHead`` reads."""

    def __init__(self, num_classes, in_channels, num_query=100, num_reg_fcs=2, transformer=None,
                 sync_cls_avg_factor=False, positional_encoding=None, loss_cls=None, loss_bbox=None,
                 loss_iou=None, train_cfg=None, test_cfg=None, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        self.num_query, self.num_classes, self.in_channels = num_query, num_classes, in_channels
        self.num_reg_fcs = num_reg_fcs
        self.sync_cls_avg_factor = sync_cls_avg_factor
        self.bg_cls_weight = 0
        sigmoid = bool((loss_cls or {}).get("use_sigmoid", False))
        self.loss_cls = types.SimpleNamespace(use_sigmoid=sigmoid)
        self.cls_out_channels = num_classes if sigmoid else num_classes + 1
        self.transformer = TRANSFORMER.build(transformer)
        self.embed_dims = self.transformer.embed_dims
        self._init_layers()


def _xavier_init(module, gain=1, bias=0, distribution="normal"):
    fn = nn.init.xavier_uniform_ if distribution == "uniform" else nn.init.xavier_normal_
    fn(module.weight, gain=gain)
    if getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)


def _constant_init(module, val, bias=0):
    nn.init.constant_(module.weight, val)
    if getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)


def _inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def _multi_apply(func, *args, **kwargs):
    results = [func(*a, **kwargs) for a in zip(*args)]
    return tuple(map(list, zip(*results)))


# ----------------------------------------------------------------------------- devkit stand-ins
class _NuScenes:
    def __init__(self, *args, **kwargs):
        pass

    def get(self, table, token):
        if table == "sample":
            chans = ["LIDAR_TOP", "RADAR_FRONT", "RADAR_FRONT_LEFT", "RADAR_FRONT_RIGHT",
                     "RADAR_BACK_LEFT", "RADAR_BACK_RIGHT"]
            return {"token": token, "data": {c: (token, c) for c in chans}}
        if table == "sample_data":
            return {"calibrated_sensor_token": token}
        if table == "calibrated_sensor":
            sample_idx, chan = token
            return {"rotation": SWEEPS[sample_idx][chan]["rotation"]}
        raise KeyError(table)


class _RadarPointCloud:
    def __init__(self, points):
        self.points = points

    @classmethod
    def from_file_multisweep(cls, nusc, sample_rec, chan, ref_chan, nsweeps=5):
        rec = SWEEPS[sample_rec["token"]][chan]
        return cls(np.array(rec["points"], dtype=np.float64)), np.array(rec["lags"], dtype=np.float64)


class _Quaternion:
    def __init__(self, rotation):
        self.rotation_matrix = np.asarray(rotation, dtype=np.float64)


# ----------------------------------------------------------------------------- installation
def _module(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
    m.__dict__.update(attrs)
    parent, _, leaf = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], leaf, m)
    return m


def _exec_reference(dotted, relpath):
    spec = importlib.util.spec_from_file_location(dotted, os.path.join(PLUGIN, relpath))
    m = importlib.util.module_from_spec(spec)
    sys.modules[dotted] = m
    spec.loader.exec_module(m)
    return m


_LOADED = {}


def load_reference(cpu=True):
    """Returns ``dict(transformer=<module>, head=<module>, coder=<module>, util=<module>)``."""
    if _LOADED:
        return _LOADED
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    ATTENTION.classes["MultiheadAttention"] = MmcvMultiheadAttention
    TRANSFORMER_LAYER.classes["DetrTransformerDecoderLayer"] = MmcvTransformerLayer
    _module("mmcv")
    _module("mmcv.cnn", xavier_init=_xavier_init, constant_init=_constant_init, Linear=nn.Linear,
            bias_init_with_prob=lambda p: float(-np.log((1 - p) / p)))
    _module("mmcv.cnn.bricks")
    _module("mmcv.cnn.bricks.registry", ATTENTION=ATTENTION, TRANSFORMER_LAYER_SEQUENCE=LAYER_SEQUENCE)
    _module("mmcv.cnn.bricks.transformer",
            MultiScaleDeformableAttention=type("MultiScaleDeformableAttention", (), {}),
            TransformerLayerSequence=MmcvLayerSequence,
            build_transformer_layer_sequence=LAYER_SEQUENCE.build)
    _module("mmcv.runner", force_fp32=lambda *a, **k: (lambda fn: fn))
    _module("mmcv.runner.base_module", BaseModule=BaseModule)
    _module("mmdet")
    _module("mmdet.core", multi_apply=_multi_apply, reduce_mean=lambda t: t)
    _module("mmdet.core.bbox", BaseBBoxCoder=object)
    _module("mmdet.core.bbox.builder", BBOX_CODERS=BBOX_CODERS)
    _module("mmdet.models", HEADS=HEADS)
    _module("mmdet.models.utils")
    _module("mmdet.models.utils.builder", TRANSFORMER=TRANSFORMER)
    _module("mmdet.models.utils.transformer", inverse_sigmoid=_inverse_sigmoid)
    _module("mmdet.models.dense_heads", DETRHead=MmdetDETRHead)
    _module("mmdet3d")
    _module("mmdet3d.core")
    _module("mmdet3d.core.bbox")
    _module("mmdet3d.core.bbox.coders", build_bbox_coder=BBOX_CODERS.build)
    _module("nuscenes")
    _module("nuscenes.nuscenes", NuScenes=_NuScenes)
    _module("nuscenes.utils")
    _module("nuscenes.utils.data_classes", RadarPointCloud=_RadarPointCloud)
    _module("pyquaternion", Quaternion=_Quaternion)
    # bypass the plugin package __init__ (it would import datasets -> real mmdet3d -> mmcv)
    for pkg in ("projects", "projects.mmdet3d_plugin", "projects.mmdet3d_plugin.core",
                "projects.mmdet3d_plugin.core.bbox"):
        _module(pkg)
    _LOADED["util"] = _exec_reference("projects.mmdet3d_plugin.core.bbox.util", "core/bbox/util.py")
    _LOADED["coder"] = _exec_reference("transcar_ref_coder", "core/bbox/coders/nms_free_coder.py")
    _LOADED["transformer"] = _exec_reference("transcar_ref_transformer", "models/utils/detr3d_transformer.py")
    if cpu:
        torch.Tensor.cuda = lambda self, *a, **k: self      # the reference hard-codes .cuda() (H:523,526,540)
    _LOADED["head"] = _exec_reference("transcar_ref_head", "models/dense_heads/detr3d_head.py")
    return _LOADED


def build_reference_head(cfg, state_dict=None):
    """Instantiate the reference ``Detr3DHead`` from a ``pts_bbox_head`` dict (eval mode)."""
    mods = load_reference()
    cfg = copy.deepcopy(cfg)
    cfg.pop("type", None)
    head = mods["head"].Detr3DHead(**cfg)
    head.init_weights()
    if state_dict is not None:
        head.load_state_dict(state_dict, strict=True)
    return head.eval()
