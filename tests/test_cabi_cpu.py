"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares,
argument validation works without a GPU, the plugin modules build from the reference config and accept reference
state-dict keys, and the product refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from transcar_b200 import _lib, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "transcar_b200.h")).read()
    return sorted(set(re.findall(r"^TC_API[^;(]*?\b(tc_\w+)\s*\(", text, flags=re.M)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 16
    assert sorted(_lib.SYMBOLS) == names, "ctypes binding and header disagree"
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.tc_abi_version() == _lib.ABI_VERSION == 8
    assert lib.tc_launch_count() >= 0


def test_struct_layouts_match_header_field_order():
    text = open(os.path.join(ROOT, "include", "transcar_b200.h")).read()
    for cname, cls in [("tc_sample_args", _lib.SampleArgs), ("tc_linear_args", _lib.LinearArgs),
                       ("tc_point_embed_args", _lib.PointEmbedArgs),
                       ("tc_attention_args", _lib.AttentionArgs),
                       ("tc_radar_geometry_args", _lib.RadarGeometryArgs), ("tc_decode_args", _lib.DecodeArgs),
                       ("tc_layernorm_args", _lib.LayerNormArgs), ("tc_layernorm_bwd_args", _lib.LayerNormBwdArgs),
                       ("tc_attention_bwd_args", _lib.AttentionBwdArgs), ("tc_sample_bwd_args", _lib.SampleBwdArgs),
                       ("tc_attention_dense_bwd_args", _lib.AttentionDenseBwdArgs), ("tc_match_cost_args", _lib.MatchCostArgs),
                       ("tc_detr_loss_args", _lib.DetrLossArgs)]:
        body = re.search(r"typedef struct \{([^}]*)\}\s*" + cname + ";", text).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                fields.append(re.sub(r"\[.*?\]", "", part.strip().split()[-1]).lstrip("*"))
        assert fields == [f[0] for f in cls._fields_], cname


def test_argument_validation_without_gpu():
    lib = _lib.load()
    assert lib.tc_linear(None, None) == -1
    assert b"NULL" in lib.tc_last_error_string()
    a = _lib.LinearArgs()
    buf = (ctypes.c_float * 16)()
    a.A = a.W = ctypes.addressof(buf)
    a.out_f32 = ctypes.addressof(buf)
    a.M, a.N, a.K, a.lda, a.ldw, a.ld_out_f32 = 0, 4, 4, 4, 4, 4
    assert lib.tc_linear(ctypes.byref(a), None) == 0             # M == 0: nothing launched
    a.N = 0
    assert lib.tc_linear(ctypes.byref(a), None) == -2
    s = _lib.SampleArgs()
    s.B, s.Q = 1, 1
    assert lib.tc_sample_fwd(ctypes.byref(s), None) == -1
    at = _lib.AttentionArgs()
    at.q = at.k = at.v = at.out = ctypes.addressof(buf)
    at.D = 64
    assert lib.tc_attention_fwd(ctypes.byref(at), None) == -2
    assert b"head dim" in lib.tc_last_error_string()
    # the fused launches validate before they touch the device: NULL blocks / pointers, unsupported sizes, misalignment
    assert lib.tc_ffn(None, None) == -1 and lib.tc_mlp(None, None) == -1
    f = _lib.FfnArgs()
    assert lib.tc_ffn(ctypes.byref(f), None) == -1 and b"required" in lib.tc_last_error_string()
    big = (ctypes.c_float * 64)()
    base = (ctypes.addressof(big) + 31) & ~31
    for name in ("X", "W1", "b1", "W2", "b2", "residual", "ln_gamma", "ln_beta", "out_f32"):
        setattr(f, name, base)
    f.M, f.C, f.H = 128, 128, 512
    assert lib.tc_ffn(ctypes.byref(f), None) == -2 and b"C = 256, H = 512" in lib.tc_last_error_string()
    f.C, f.ldx, f.ldw1, f.ldw2, f.ld_residual, f.ld_out_f32 = 256, 512, 512, 1024, 256, 256
    f.X = base + 2
    assert lib.tc_ffn(ctypes.byref(f), None) == -3 and b"16-byte aligned" in lib.tc_last_error_string()
    m = _lib.MlpArgs()
    for name in ("X", "W1", "b1", "W2", "b2", "W3", "b3", "out_f32"):
        setattr(m, name, base)
    m.M, m.C, m.N3 = 128, 256, 48
    assert lib.tc_mlp(ctypes.byref(m), None) == -2 and b"N3 <= 32" in lib.tc_last_error_string()
    m.N3, m.ldx, m.ldw1, m.ldw2, m.ldw3, m.ld_out_f32 = 10, 512, 512, 512, 512, 10
    m.ln1_gamma = base                                                    # gamma without beta
    assert lib.tc_mlp(ctypes.byref(m), None) == -1 and b"pairs" in lib.tc_last_error_string()
    m.ln1_gamma, m.tail = None, _lib.TC_TAIL_BOX                          # a tail without its input
    assert lib.tc_mlp(ctypes.byref(m), None) == -1 and b"tail" in lib.tc_last_error_string()


def test_plugin_builds_from_reference_config_and_loads_reference_keys():
    from transcar_b200 import plugin
    head = plugin.build_head(synthetic.head_config(900))
    sd = synthetic.make_state_dict(0, 900)
    res = head.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert sum(p.numel() for p in head.parameters()) == 8728385
    assert [n for n, *_ in synthetic.state_dict_spec()] == list(sd.keys())
    for reg, name in [(plugin.ATTENTION, "Detr3DCrossAtten"), (plugin.TRANSFORMER, "Detr3DTransformer"),
                      (plugin.TRANSFORMER_LAYER_SEQUENCE, "Detr3DTransformerDecoder"), (plugin.HEADS, "Detr3DHead"),
                      (plugin.BBOX_CODERS, "NMSFreeCoder")]:
        assert name in reg.module_dict


def test_no_cpu_fallback():
    from transcar_b200 import ops, plugin
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.linear(torch.zeros(2, 4), torch.zeros(3, 4))
    head = plugin.build_head(synthetic.head_config(32)).eval()
    feats = synthetic.make_feats(0, 1, "tiny")
    metas = synthetic.make_img_metas(1)
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        head(feats, metas)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "transcar_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
    bench = open(os.path.join(ROOT, "bench.py")).read() if os.path.exists(os.path.join(ROOT, "bench.py")) else ""
    # bench.py may touch the oracle only inside its cpu-baseline / reference-arm functions
    for m in re.finditer(r"^(\s*)(from|import)\s+oracle\b", bench, flags=re.M):
        assert len(m.group(1)) > 0, "oracle import at module level of bench.py"


def test_synthetic_workload_statistics():
    """Real calibration => ~18 % of (query, camera) pairs are valid; radar ~1.5k points."""
    calib = synthetic.load_calibration()
    assert calib.shape == (2, 6, 4, 4)
    metas = synthetic.make_img_metas(2, seed=0)
    n = [m["radar_tokens"].shape[0] for m in metas]
    assert all(1300 < x < 1600 for x in n) and metas[0]["radar_tokens"].shape[1] == 36
    f = synthetic.make_feats(0, 1, "tiny", channels_last=True)[0]
    assert f.shape == (1, 6, 256, 16, 28) and f.stride(2) == 1
