"""ctypes binding of ``libtranscar_b200.so`` (the C ABI declared in ``include/transcar_b200.h``).

There is no fallback: if the shared library is missing this module raises at first use with the
build command, and every call raises ``RuntimeError`` carrying ``tc_last_error_string()`` on a
non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtranscar_b200.so")

TC_F32, TC_BF16, TC_BF16X2, TC_F16 = 0, 1, 2, 3
TC_MAX_LEVELS, TC_MAX_CAMS = 4, 8
ABI_VERSION = 8
TC_SAMPLE_ALL_CAMS, TC_SAMPLE_WEIGHTS_GIVEN = 1, 2
TC_TAIL_NONE, TC_TAIL_REF_UPDATE, TC_TAIL_BOX = 0, 1, 2
TC_ATTN_AUTO, TC_ATTN_TENSOR, TC_ATTN_SIMT, TC_ATTN_SPARSE = 0, 1, 2, 3

_vp, _i32, _i64, _f32, _u8p = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_void_p


class SampleArgs(C.Structure):
    _fields_ = [("feat", _vp * TC_MAX_LEVELS), ("H", _i32 * TC_MAX_LEVELS), ("W", _i32 * TC_MAX_LEVELS),
                ("num_levels", _i32), ("B", _i32), ("N", _i32), ("Q", _i32), ("C", _i32),
                ("feat_dtype", _i32), ("out_dtype", _i32),
                ("ref", _vp), ("lidar2img", _vp), ("attn_logits", _vp),
                ("pc_range", _f32 * 6), ("img_w", _f32), ("img_h", _f32),
                ("out", _vp), ("mask", _vp), ("flags", _i32)]


class LinearArgs(C.Structure):
    _fields_ = [("A", _vp), ("a_dtype", _i32), ("lda", _i64),
                ("W", _vp), ("w_dtype", _i32), ("ldw", _i64),
                ("M", _i32), ("N", _i32), ("K", _i32),
                ("bias", _vp),
                ("row_bias", _vp), ("row_bias_period", _i32), ("ld_row_bias", _i64),
                ("row_gate", _vp),
                ("residual", _vp), ("ld_residual", _i64),
                ("residual2", _vp), ("ld_residual2", _i64),
                ("ln_gamma", _vp), ("ln_beta", _vp), ("ln_eps", _f32),
                ("relu", _i32),
                ("post_add", _vp), ("ld_post_add", _i64),
                ("out_f32", _vp), ("ld_out_f32", _i64),
                ("out_bf16", _vp), ("ld_out_bf16", _i64),
                ("out16_dtype", _i32),
                ("tail", _i32), ("tail_in", _vp), ("ld_tail_in", _i64), ("tail_ref_out", _vp), ("tail_geom_out", _vp),
                ("tail_xy_col", _i32), ("tail_z_col", _i32), ("tail_from_norm", _i32),
                ("tail_pc_range", _f32 * 6), ("tail_r_lo", _f32), ("tail_r_hi", _f32), ("w_static", _i32)]


class FfnArgs(C.Structure):
    _fields_ = [("X", _vp), ("ldx", _i64), ("W1", _vp), ("ldw1", _i64), ("b1", _vp), ("W2", _vp), ("ldw2", _i64), ("b2", _vp),
                ("residual", _vp), ("ld_residual", _i64), ("ln_gamma", _vp), ("ln_beta", _vp), ("ln_eps", _f32),
                ("out_f32", _vp), ("ld_out_f32", _i64), ("out16", _vp), ("ld_out16", _i64),
                ("M", _i32), ("C", _i32), ("H", _i32), ("w_static", _i32)]


class MlpArgs(C.Structure):
    _fields_ = [("X", _vp), ("ldx", _i64),
                ("W1", _vp), ("ldw1", _i64), ("b1", _vp), ("ln1_gamma", _vp), ("ln1_beta", _vp),
                ("W2", _vp), ("ldw2", _i64), ("b2", _vp), ("ln2_gamma", _vp), ("ln2_beta", _vp),
                ("W3", _vp), ("ldw3", _i64), ("b3", _vp),
                ("ln_eps", _f32), ("out_f32", _vp), ("ld_out_f32", _i64),
                ("M", _i32), ("C", _i32), ("N3", _i32), ("w_static", _i32),
                ("tail", _i32), ("tail_in", _vp), ("ld_tail_in", _i64), ("tail_ref_out", _vp), ("tail_geom_out", _vp),
                ("tail_xy_col", _i32), ("tail_z_col", _i32), ("tail_from_norm", _i32),
                ("tail_pc_range", _f32 * 6), ("tail_r_lo", _f32), ("tail_r_hi", _f32)]


class PointEmbedArgs(C.Structure):
    _fields_ = [("x", _vp), ("ldx", _i64), ("M", _i32), ("C", _i32), ("logit_input", _i32),
                ("weight", _vp), ("bias", _vp), ("ln_gamma", _vp), ("ln_beta", _vp), ("ln_eps", _f32),
                ("out_f32", _vp), ("out_bf16", _vp), ("out16_dtype", _i32)]


class AttentionArgs(C.Structure):
    _fields_ = [("q", _vp), ("k", _vp), ("v", _vp),
                ("ldq", _i64), ("ldk", _i64), ("ldv", _i64),
                ("q_batch_stride", _i64), ("k_batch_stride", _i64), ("v_batch_stride", _i64),
                ("qkv_dtype", _i32),
                ("B", _i32), ("Lq", _i32), ("Lk", _i32), ("heads", _i32), ("D", _i32),
                ("scale", _f32),
                ("geom", _vp), ("key_xy", _vp),
                ("out", _vp), ("ldo", _i64), ("out_dtype", _i32),
                ("row_any", _vp), ("algo", _i32),
                ("dropout_p", _f32), ("dropout_seed", C.c_uint64), ("dropout_stream", C.c_uint64),
                ("attn_blocked", _vp), ("key_blocked", _vp)]


class RadarGeometryArgs(C.Structure):
    _fields_ = [("centre", _vp), ("ld_centre", _i64), ("centre_is_normalised", _i32),
                ("code", _vp), ("ld_code", _i64),
                ("M", _i32),
                ("pc_range", _f32 * 6), ("r_lo", _f32), ("r_hi", _f32),
                ("geom", _vp)]


class DecodeArgs(C.Structure):
    _fields_ = [("cls", _vp), ("code", _vp),
                ("B", _i32), ("Q", _i32), ("classes", _i32), ("max_num", _i32),
                ("post_center_range", _f32 * 6),
                ("boxes", _vp), ("scores", _vp), ("labels", _vp), ("keep", _vp),
                ("workspace", _vp), ("records", _vp)]


class LayerNormArgs(C.Structure):
    _fields_ = [("x", _vp), ("ldx", _i64), ("M", _i32), ("N", _i32),
                ("gamma", _vp), ("beta", _vp), ("eps", _f32), ("relu", _i32),
                ("y_f32", _vp), ("y_bf16", _vp), ("ldy", _i64),
                ("mean", _vp), ("rstd", _vp)]


class LayerNormBwdArgs(C.Structure):
    _fields_ = [("dy", _vp), ("ld_dy", _i64), ("x", _vp), ("ldx", _i64),
                ("mean", _vp), ("rstd", _vp), ("gamma", _vp),
                ("M", _i32), ("N", _i32),
                ("add", _vp), ("ld_add", _i64),
                ("dx", _vp), ("ld_dx", _i64),
                ("dgamma", _vp), ("dbeta", _vp)]


class AttentionBwdArgs(C.Structure):
    _fields_ = [("q", _vp), ("k", _vp), ("v", _vp), ("dout", _vp),
                ("ldq", _i64), ("ldk", _i64), ("ldv", _i64), ("ld_dout", _i64),
                ("q_batch_stride", _i64), ("k_batch_stride", _i64), ("v_batch_stride", _i64),
                ("B", _i32), ("Lq", _i32), ("Lk", _i32), ("heads", _i32), ("D", _i32),
                ("scale", _f32),
                ("geom", _vp), ("key_xy", _vp),
                ("dq", _vp), ("ld_dq", _i64),
                ("dk", _vp), ("dv", _vp), ("ld_dk", _i64), ("ld_dv", _i64), ("dk_batch_stride", _i64), ("dv_batch_stride", _i64),
                ("dropout_p", _f32), ("dropout_seed", C.c_uint64), ("dropout_stream", C.c_uint64)]


class SampleBwdArgs(C.Structure):
    _fields_ = [("feat", _vp * TC_MAX_LEVELS), ("H", _i32 * TC_MAX_LEVELS), ("W", _i32 * TC_MAX_LEVELS),
                ("num_levels", _i32), ("B", _i32), ("N", _i32), ("Q", _i32), ("C", _i32),
                ("feat_dtype", _i32),
                ("ref", _vp), ("lidar2img", _vp), ("attn_logits", _vp),
                ("pc_range", _f32 * 6), ("img_w", _f32), ("img_h", _f32),
                ("dout", _vp), ("d_feat", _vp * TC_MAX_LEVELS), ("d_logits", _vp), ("d_ref", _vp)]


class AttentionDenseBwdArgs(C.Structure):
    _fields_ = [("q", _vp), ("k", _vp), ("v", _vp), ("o", _vp), ("dout", _vp),
                ("ldq", _i64), ("ldk", _i64), ("ldv", _i64), ("ldo", _i64), ("ld_dout", _i64),
                ("q_batch_stride", _i64), ("k_batch_stride", _i64), ("v_batch_stride", _i64), ("o_batch_stride", _i64),
                ("dout_batch_stride", _i64),
                ("B", _i32), ("Lq", _i32), ("Lk", _i32), ("heads", _i32), ("D", _i32),
                ("scale", _f32),
                ("dq", _vp), ("dk", _vp), ("dv", _vp),
                ("workspace", _vp),
                ("dropout_p", _f32), ("dropout_seed", C.c_uint64), ("dropout_stream", C.c_uint64)]


class MatchCostArgs(C.Structure):
    _fields_ = [("cls", _vp), ("bbox", _vp), ("gt_boxes", _vp), ("gt_labels", _vp), ("gt_offsets", _vp),
                ("layers", _i32), ("B", _i32), ("Q", _i32), ("classes", _i32), ("Gmax", _i32),
                ("cls_weight", _f32), ("reg_weight", _f32), ("alpha", _f32), ("gamma", _f32), ("eps", _f32),
                ("cost", _vp)]


class DetrLossArgs(C.Structure):
    _fields_ = [("cls", _vp), ("bbox", _vp), ("assigned", _vp), ("gt_boxes", _vp), ("gt_labels", _vp),
                ("code_weights", _vp), ("cls_avg", _vp), ("pos_avg", _vp),
                ("layers", _i32), ("B", _i32), ("Q", _i32), ("classes", _i32),
                ("alpha", _f32), ("gamma", _f32), ("loss_cls_weight", _f32), ("loss_bbox_weight", _f32),
                ("loss_cls", _vp), ("loss_bbox", _vp), ("d_cls", _vp), ("d_bbox", _vp)]


# every symbol include/transcar_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "tc_abi_version": (C.c_int, []),
    "tc_last_error_string": (C.c_char_p, []),
    "tc_check_device": (C.c_int, []),
    "tc_launch_count": (C.c_uint64, []),
    "tc_debug_trace": (C.c_int, [_vp, _i64]),
    "tc_sample_fwd": (C.c_int, [C.POINTER(SampleArgs), _vp]),
    "tc_nchw_to_nhwc": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "tc_linear": (C.c_int, [C.POINTER(LinearArgs), _vp]),
    "tc_ffn": (C.c_int, [C.POINTER(FfnArgs), _vp]),
    "tc_mlp": (C.c_int, [C.POINTER(MlpArgs), _vp]),
    "tc_point_embed": (C.c_int, [C.POINTER(PointEmbedArgs), _vp]),
    "tc_attention_fwd": (C.c_int, [C.POINTER(AttentionArgs), _vp]),
    "tc_radar_geometry": (C.c_int, [C.POINTER(RadarGeometryArgs), _vp]),
    "tc_radar_mask": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "tc_ref_update": (C.c_int, [_vp, _i64, _vp, _vp, _i32, _vp]),
    "tc_box_anchor_add": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, C.POINTER(_f32 * 6), _i32, _vp]),
    "tc_cast_bf16": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _vp]),
    "tc_cast_split": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _vp]),
    "tc_decode_workspace_bytes": (C.c_int64, [_i32, _i32, _i32]),
    "tc_decode": (C.c_int, [C.POINTER(DecodeArgs), _vp]),
    "tc_transpose": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _i64, _i32, _i32, _vp]),
    "tc_colsum": (C.c_int, [_vp, _i32, _i64, _i32, _i32, _vp, _vp]),
    "tc_layernorm_fwd": (C.c_int, [C.POINTER(LayerNormArgs), _vp]),
    "tc_layernorm_bwd": (C.c_int, [C.POINTER(LayerNormBwdArgs), _vp]),
    "tc_mask_grad": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _vp]),
    "tc_attention_sparse_bwd": (C.c_int, [C.POINTER(AttentionBwdArgs), _vp]),
    "tc_match_cost": (C.c_int, [C.POINTER(MatchCostArgs), _vp]),
    "tc_detr_loss": (C.c_int, [C.POINTER(DetrLossArgs), _vp]),
    "tc_sample_bwd": (C.c_int, [C.POINTER(SampleBwdArgs), _vp]),
    "tc_attention_dense_bwd": (C.c_int, [C.POINTER(AttentionDenseBwdArgs), _vp]),
    "tc_pointwise": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp]),
    "tc_dropout": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _f32, C.c_uint64, C.c_uint64, _vp]),
    "tc_add_rows": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "tc_period_sum": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp]),
}

_lib = None


def load():
    """Load the shared library once; raise (never fall back) when it is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"transcar_b200: CUDA library {LIB_PATH} is missing and there is no CPU fallback. "
            f"Build it with `python -m transcar_b200.build` (nvcc, sm_100a).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)       # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.tc_abi_version() != ABI_VERSION:
        raise RuntimeError(f"transcar_b200: ABI version mismatch (library {lib.tc_abi_version()}, binding {ABI_VERSION})")
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        msg = load().tc_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"transcar_b200.{what} failed with code {code}: {msg}")


def launch_count():
    return int(load().tc_launch_count())
