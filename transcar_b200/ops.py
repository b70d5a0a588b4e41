"""Torch-tensor front end of the C ABI: PyTorch owns memory and streams, the library does the math.

Every function enqueues on ``torch.cuda.current_stream()`` and returns immediately.  Inputs must be
CUDA tensors; nothing here computes with torch ops and nothing falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from ._lib import TC_BF16, TC_BF16X2, TC_F16, TC_F32

_DT = {torch.float32: TC_F32, torch.bfloat16: TC_BF16, torch.float16: TC_F16}


class SplitBf16:
    """Split-bf16 matrix (``TC_BF16X2``): logical ``[..., cols]`` stored as bf16 ``[..., 2 * cols]`` = ``hi | lo`` with
    ``hi = bf16(x)``, ``lo = bf16(x - hi)``.  The operand format of the bf16x3 tensor-core mode."""
    __slots__ = ("t", "static")

    def __init__(self, t):
        assert t.dtype == torch.bfloat16 and t.shape[-1] % 2 == 0
        self.t = t
        self.static = False          # see mark_static()

    @property
    def shape(self):
        return (*self.t.shape[:-1], self.t.shape[-1] // 2)

    @property
    def device(self):
        return self.t.device

    def view(self, *shape):
        """View with the LOGICAL trailing dimension, e.g. ``[B,Q,C] -> view(B*Q, C)``."""
        return SplitBf16(self.t.view(*shape[:-1], 2 * shape[-1]))

    def float(self):
        c = self.t.shape[-1] // 2
        return self.t[..., :c].float() + self.t[..., c:].float()


def _out16(kind, shape, device):
    """Allocate the 16-bit output of a kernel: kind in {'bf16', 'split', 'f16'} -> (tensor, wrapped, dtype code)."""
    if kind == "split":
        t = torch.empty((*shape[:-1], 2 * shape[-1]), device=device, dtype=torch.bfloat16)
        return t, SplitBf16(t), TC_BF16X2
    if kind == "f16":
        t = torch.empty(shape, device=device, dtype=torch.float16)
        return t, t, TC_F16
    if kind != "bf16":
        raise ValueError(f"unknown 16-bit output kind {kind!r}")
    t = torch.empty(shape, device=device, dtype=torch.bfloat16)
    return t, t, TC_BF16

# Profiling hook (tools/step_timeline.py): when set to a list, every library call is bracketed by a pair of CUDA
# events (external=True, so the pair becomes two nodes of a captured CUDA graph) and (label, start, end) is appended.
TIMELINE = None
TIMELINE_EXTERNAL = True


def _call(label, fn, *args):
    if TIMELINE is None:
        return fn(*args)
    e0 = torch.cuda.Event(enable_timing=True, external=TIMELINE_EXTERNAL)
    e1 = torch.cuda.Event(enable_timing=True, external=TIMELINE_EXTERNAL)
    e0.record()
    rc = fn(*args)
    e1.record()
    TIMELINE.append((label, e0, e1))
    return rc


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _need(t, name, dtype=None, last_contig=True):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"transcar_b200: `{name}` must be a CUDA tensor (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"transcar_b200: `{name}` must be {dtype}, got {t.dtype}")
    if last_contig and t.dim() > 0 and t.stride(-1) != 1 and t.shape[-1] != 1:
        raise RuntimeError(f"transcar_b200: `{name}` must be contiguous in its last dimension")
    return t


def _rows(t, name):
    """2-D view [M, N] with a uniform row stride; returns (tensor, ld)."""
    if t.dim() != 2:
        t = t.reshape(-1, t.shape[-1])
    _need(t, name)
    return t, t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


# --------------------------------------------------------------------------------------- K1
def is_channels_last_5d(f):
    """logical [B,N,C,H,W] stored as [B,N,H,W,C]"""
    B, N, Cc, H, W = f.shape
    return f.stride() == (N * H * W * Cc, H * W * Cc, 1, W * Cc, Cc)


def to_channels_last(f, dtype=None):
    """[B,N,C,H,W] in any layout -> same logical tensor stored channels-last, via tc_nchw_to_nhwc when a
    physical re-layout is needed (fp32 NCHW input).  Zero-copy when already channels-last."""
    dtype = dtype or f.dtype
    if is_channels_last_5d(f) and f.dtype == dtype:
        return f
    B, N, Cc, H, W = f.shape
    if f.dtype != torch.float32 or not f.is_contiguous():
        raise RuntimeError("transcar_b200: feature maps must be channels-last (any dtype) or contiguous NCHW fp32")
    _need(f, "feat")
    out = torch.empty((B, N, H, W, Cc), device=f.device, dtype=dtype)
    lib = _lib.load()
    _lib.check(_call("nchw_to_nhwc", lib.tc_nchw_to_nhwc, _ptr(f), _ptr(out), _DT[dtype], B * N, Cc, H * W, _stream()), "nchw_to_nhwc")
    return out.permute(0, 1, 4, 2, 3)


def sample_fwd(feats, ref, lidar2img, attn_logits, pc_range, img_w, img_h, out_dtype=torch.float32,
               want_mask=False, out=None, all_cams=False, weights_given=False):
    """feats: 4 x logical [B,N,C,H,W] channels-last; ref [B,Q,3]; lidar2img [B,N,4,4]; attn_logits [B,Q,N*L].
    ``out_dtype``: torch.float32 / torch.bfloat16 / ``"split"`` (SplitBf16 [B,Q,C]).
    Returns (out [B,Q,C], mask [B,Q,N] uint8 or None)."""
    lib = _lib.load()
    a = _lib.SampleArgs()
    B, N, Cc = feats[0].shape[:3]
    Q = ref.shape[1]
    if len(feats) != 4:
        raise RuntimeError("transcar_b200.sample_fwd: exactly 4 feature levels are supported")
    for l, f in enumerate(feats):
        _need(f, f"feats[{l}]", last_contig=False)
        if not is_channels_last_5d(f):
            raise RuntimeError(f"transcar_b200.sample_fwd: feats[{l}] is not channels-last; use ops.to_channels_last")
        if f.dtype != feats[0].dtype or f.shape[:3] != feats[0].shape[:3]:
            raise RuntimeError("transcar_b200.sample_fwd: feature levels disagree in dtype / batch / cams / channels")
        a.feat[l] = f.data_ptr()
        a.H[l], a.W[l] = f.shape[3], f.shape[4]
    ref = _need(ref, "ref", torch.float32).contiguous()
    lidar2img = _need(lidar2img, "lidar2img", torch.float32).contiguous()
    attn_logits = _need(attn_logits, "attn_logits", torch.float32).contiguous()
    assert ref.shape == (B, Q, 3) and lidar2img.shape == (B, N, 4, 4) and attn_logits.shape == (B, Q, N * 4)
    ret = out
    if out is None:
        if out_dtype == "split":
            out, ret, _ = _out16("split", (B, Q, Cc), ref.device)
        else:
            ret = out = torch.empty((B, Q, Cc), device=ref.device, dtype=out_dtype)
    elif isinstance(out, SplitBf16):
        out = out.t
    mask = torch.empty((B, Q, N), device=ref.device, dtype=torch.uint8) if want_mask else None
    a.num_levels, a.B, a.N, a.Q, a.C = 4, B, N, Q, Cc
    a.feat_dtype = _DT[feats[0].dtype]
    a.out_dtype = TC_BF16X2 if isinstance(ret, SplitBf16) else _DT[out.dtype]
    a.ref, a.lidar2img, a.attn_logits = ref.data_ptr(), lidar2img.data_ptr(), attn_logits.data_ptr()
    for i in range(6):
        a.pc_range[i] = float(pc_range[i])
    a.img_w, a.img_h = float(img_w), float(img_h)
    a.out = out.data_ptr()
    a.mask = mask.data_ptr() if mask is not None else None
    a.flags = (_lib.TC_SAMPLE_ALL_CAMS if all_cams else 0) | (_lib.TC_SAMPLE_WEIGHTS_GIVEN if weights_given else 0)
    _lib.check(_call("sample", lib.tc_sample_fwd, C.byref(a), _stream()), "sample_fwd")
    return ret, mask


# --------------------------------------------------------------------------------------- K3
def mark_static(w):
    """Tag a GEMM weight operand as a parameter that no launch on the stream writes (``tc_linear_args.w_static``): the
    tensor-core Linear then prefetches its first W tiles before it waits for the previous kernel.  Never tag an activation."""
    if isinstance(w, SplitBf16):
        w.static = True
    else:
        w._tc_static = True
    return w


def linear(A, W, bias=None, *, row_bias=None, row_bias_period=0, row_gate=None, residual=None, residual2=None,
           ln=None, ln_eps=1e-5, relu=False, post_add=None, out_f32=None, out_bf16=None,
           want_f32=True, want_bf16=False, out16="bf16", tail=None):
    """Y = epilogue(A @ W^T) - see ``tc_linear`` in include/transcar_b200.h.  A [M,K], W [N,K]; both fp32, both bf16
    or both :class:`SplitBf16` (bf16x3).  ``ln`` = (gamma, beta).  ``out16`` = format of the 16-bit output when
    ``want_bf16``: 'bf16', 'split' (SplitBf16) or 'f16'.  ``tail`` (N <= 32): a fused row-local stage,
    ``dict(kind='ref_update', ref=[M,3], pc_range=, geom=(r_lo, r_hi) | None)`` or
    ``dict(kind='box', anchor=[M,ld], xy_col=, z_col=, from_norm=, pc_range=, geom=(r_lo, r_hi) | None)``; its outputs are
    returned in ``tail['ref_out']`` / ``tail['geom_out']``.  Returns (out_f32 or None, 16-bit output or None)."""
    lib = _lib.load()
    w_static = bool(getattr(W, "static", False) or getattr(W, "_tc_static", False))
    split = isinstance(A, SplitBf16)
    if split != isinstance(W, SplitBf16):
        raise RuntimeError("transcar_b200.linear: split-bf16 operands come in pairs (A and W)")
    if split:
        A, W = A.t, W.t
    A, lda = _rows(_need(A, "A"), "A")
    W, ldw = _rows(_need(W, "W"), "W")
    M, K = A.shape
    N = W.shape[0]
    assert W.shape[1] == K, (A.shape, W.shape)
    if split:
        K //= 2
    a = _lib.LinearArgs()
    a.A, a.a_dtype, a.lda = A.data_ptr(), TC_BF16X2 if split else _DT[A.dtype], lda
    a.W, a.w_dtype, a.ldw = W.data_ptr(), TC_BF16X2 if split else _DT[W.dtype], ldw
    a.M, a.N, a.K = M, N, K
    keep = [A, W]
    if bias is not None:
        a.bias = _need(bias, "bias", torch.float32).data_ptr()
    if row_bias is not None:
        rb, ldrb = _rows(_need(row_bias, "row_bias", torch.float32), "row_bias")
        a.row_bias, a.ld_row_bias = rb.data_ptr(), ldrb
        a.row_bias_period = row_bias_period or rb.shape[0]
        keep.append(rb)
    if row_gate is not None:
        a.row_gate = _need(row_gate, "row_gate", torch.uint8).data_ptr()
    if residual is not None:
        r, ldr = _rows(_need(residual, "residual", torch.float32), "residual")
        a.residual, a.ld_residual = r.data_ptr(), ldr
        keep.append(r)
    if residual2 is not None:
        r2, ldr2 = _rows(_need(residual2, "residual2", torch.float32), "residual2")
        a.residual2, a.ld_residual2 = r2.data_ptr(), ldr2
        keep.append(r2)
    if ln is not None:
        a.ln_gamma = _need(ln[0], "ln_gamma", torch.float32).data_ptr()
        a.ln_beta = _need(ln[1], "ln_beta", torch.float32).data_ptr()
        a.ln_eps = float(ln_eps)
    a.relu = 1 if relu else 0
    a.w_static = 1 if w_static else 0
    if post_add is not None:
        pa, ldpa = _rows(_need(post_add, "post_add", torch.float32), "post_add")
        a.post_add, a.ld_post_add = pa.data_ptr(), ldpa
        keep.append(pa)
    if out_f32 is None and want_f32:
        out_f32 = torch.empty((M, N), device=A.device, dtype=torch.float32)
    ret16 = out_bf16
    if out_bf16 is None and want_bf16:
        out_bf16, ret16, a.out16_dtype = _out16(out16, (M, N), A.device)
    elif out_bf16 is not None:
        if isinstance(out_bf16, SplitBf16):
            out_bf16, a.out16_dtype = out_bf16.t, TC_BF16X2
        else:
            a.out16_dtype = _DT[out_bf16.dtype]
    if out_f32 is not None:
        o, ldo = _rows(_need(out_f32, "out_f32", torch.float32), "out_f32")
        if o.shape != (M, N):
            raise RuntimeError(f"transcar_b200.linear: out_f32 has shape {tuple(o.shape)}, expected {(M, N)}")
        a.out_f32, a.ld_out_f32 = o.data_ptr(), ldo
    if out_bf16 is not None:
        o, ldo = _rows(_need(out_bf16, "out16"), "out16")
        if o.shape != (M, 2 * N if a.out16_dtype == TC_BF16X2 else N):
            raise RuntimeError(f"transcar_b200.linear: 16-bit output has shape {tuple(o.shape)} for M, N = {(M, N)}")
        a.out_bf16, a.ld_out_bf16 = o.data_ptr(), ldo
    if tail is not None:
        _fill_tail(a, tail, M, A.device, keep)
    label = "linear" if TIMELINE is None else (
        f"linear M{M} N{N} K{K} {'bf16x3' if split else 'bf16' if A.dtype == torch.bfloat16 else 'f32'}" + ("+rb" if row_bias is not None else "") +
        ("+gate" if row_gate is not None else "") + ("+res" if residual is not None else "") +
        ("+res2" if residual2 is not None else "") + ("+ln" if ln is not None else "") + ("+relu" if relu else "") +
        ("+post" if post_add is not None else "") + (f"+tail:{tail['kind']}" if tail is not None else ""))
    _lib.check(_call(label, lib.tc_linear, C.byref(a), _stream()), "linear")
    return out_f32, ret16


def ffn_supported(x, W1, W2):
    """True when ``tc_ffn`` (the fused feed-forward launch) covers these operands: split bf16, C = 256, H = 512."""
    return (isinstance(x, SplitBf16) and isinstance(W1, SplitBf16) and isinstance(W2, SplitBf16)
            and x.shape[-1] == 256 and tuple(W1.shape) == (512, 256) and tuple(W2.shape) == (256, 512))


def ffn(x, W1, b1, W2, b2, residual, ln, ln_eps=1e-5, want_f32=True, want_16=True):
    """LayerNorm(residual + W2 relu(W1 x + b1) + b2) as ONE launch (``tc_ffn``, csrc/ffn_tc.cu).  x [M, C], W1 [H, C],
    W2 [C, H] are :class:`SplitBf16`; residual fp32 [M, C]; ``ln`` = (gamma, beta).  Returns (fp32 [M, C] or None,
    SplitBf16 [M, C] or None)."""
    lib = _lib.load()
    if not ffn_supported(x, W1, W2):
        raise RuntimeError("transcar_b200.ffn: needs split-bf16 operands with C = 256, H = 512 (use two linear() calls)")
    w_static = bool(W1.static and W2.static)
    X, ldx = _rows(_need(x.t, "x"), "x")
    w1, ldw1 = _rows(_need(W1.t, "W1"), "W1")
    w2, ldw2 = _rows(_need(W2.t, "W2"), "W2")
    M, Cc, H = X.shape[0], X.shape[1] // 2, w1.shape[0]
    r, ldr = _rows(_need(residual, "residual", torch.float32), "residual")
    if r.shape != (M, Cc):
        raise RuntimeError(f"transcar_b200.ffn: residual has shape {tuple(r.shape)}, expected {(M, Cc)}")
    a = _lib.FfnArgs()
    a.X, a.ldx, a.W1, a.ldw1, a.W2, a.ldw2 = X.data_ptr(), ldx, w1.data_ptr(), ldw1, w2.data_ptr(), ldw2
    a.b1 = _need(b1, "b1", torch.float32).data_ptr()
    a.b2 = _need(b2, "b2", torch.float32).data_ptr()
    a.residual, a.ld_residual = r.data_ptr(), ldr
    a.ln_gamma = _need(ln[0], "ln_gamma", torch.float32).data_ptr()
    a.ln_beta = _need(ln[1], "ln_beta", torch.float32).data_ptr()
    a.ln_eps = float(ln_eps)
    a.M, a.C, a.H, a.w_static = M, Cc, H, 1 if w_static else 0
    o32 = torch.empty((M, Cc), device=X.device, dtype=torch.float32) if want_f32 else None
    o16 = SplitBf16(torch.empty((M, 2 * Cc), device=X.device, dtype=torch.bfloat16)) if want_16 else None
    if o32 is not None:
        a.out_f32, a.ld_out_f32 = o32.data_ptr(), Cc
    if o16 is not None:
        a.out16, a.ld_out16 = o16.t.data_ptr(), 2 * Cc
    label = "ffn" if TIMELINE is None else f"ffn M{M} C{Cc} H{H} bf16x3+res+ln"
    _lib.check(_call(label, lib.tc_ffn, C.byref(a), _stream()), "ffn")
    return o32, o16


def mlp_supported(x, W1, W2, W3):
    """True when ``tc_mlp`` (the fused three-layer head) covers these operands: split bf16, C = 256, at most 32 outputs."""
    return (all(isinstance(t, SplitBf16) for t in (x, W1, W2, W3)) and x.shape[-1] == 256
            and tuple(W1.shape) == (256, 256) and tuple(W2.shape) == (256, 256) and W3.shape[1] == 256 and W3.shape[0] <= 32)


def _fill_tail(a, tail, M, device, keep):
    """Shared by linear() and mlp(): the row-local tail fields of the argument block (see ``tc_linear``)."""
    tin, ldt = _rows(_need(tail["ref" if tail["kind"] == "ref_update" else "anchor"], "tail_in", torch.float32), "tail_in")
    keep.append(tin)
    a.tail = _lib.TC_TAIL_REF_UPDATE if tail["kind"] == "ref_update" else _lib.TC_TAIL_BOX
    a.tail_in, a.ld_tail_in = tin.data_ptr(), ldt
    for i in range(6):
        a.tail_pc_range[i] = float(tail["pc_range"][i])
    if tail["kind"] == "ref_update":
        tail["ref_out"] = torch.empty((M, 3), device=device, dtype=torch.float32)
        a.tail_ref_out = tail["ref_out"].data_ptr()
    else:
        a.tail_xy_col, a.tail_z_col, a.tail_from_norm = tail["xy_col"], tail["z_col"], 1 if tail["from_norm"] else 0
    if tail.get("geom") is not None:
        tail["geom_out"] = torch.empty((M, 8), device=device, dtype=torch.float32)
        a.tail_geom_out = tail["geom_out"].data_ptr()
        a.tail_r_lo, a.tail_r_hi = float(tail["geom"][0]), float(tail["geom"][1])


def mlp(x, W1, b1, W2, b2, W3, b3, ln1=None, ln2=None, ln_eps=1e-5, out_f32=None, tail=None):
    """W3 f2(W2 f1(W1 x + b1) + b2) + b3 as ONE launch (``tc_mlp``, csrc/mlp_tc.cu); f = ReLU, or ReLU(LayerNorm(.)) when
    ``ln1`` / ``ln2`` = (gamma, beta) is given.  x [M, 256], W1 / W2 [256, 256], W3 [N3 <= 32, 256] are :class:`SplitBf16`.
    ``tail`` as in :func:`linear`.  Returns the fp32 [M, N3] output."""
    lib = _lib.load()
    if not mlp_supported(x, W1, W2, W3):
        raise RuntimeError("transcar_b200.mlp: needs split-bf16 operands with C = 256 and N3 <= 32 (use three linear() calls)")
    X, ldx = _rows(_need(x.t, "x"), "x")
    w1, ldw1 = _rows(_need(W1.t, "W1"), "W1")
    w2, ldw2 = _rows(_need(W2.t, "W2"), "W2")
    w3, ldw3 = _rows(_need(W3.t, "W3"), "W3")
    M, N3 = X.shape[0], w3.shape[0]
    a = _lib.MlpArgs()
    a.X, a.ldx, a.W1, a.ldw1, a.W2, a.ldw2, a.W3, a.ldw3 = (X.data_ptr(), ldx, w1.data_ptr(), ldw1, w2.data_ptr(), ldw2,
                                                           w3.data_ptr(), ldw3)
    a.b1 = _need(b1, "b1", torch.float32).data_ptr()
    a.b2 = _need(b2, "b2", torch.float32).data_ptr()
    a.b3 = _need(b3, "b3", torch.float32).data_ptr()
    if ln1 is not None:
        a.ln1_gamma, a.ln1_beta = _need(ln1[0], "ln1_gamma", torch.float32).data_ptr(), _need(ln1[1], "ln1_beta", torch.float32).data_ptr()
    if ln2 is not None:
        a.ln2_gamma, a.ln2_beta = _need(ln2[0], "ln2_gamma", torch.float32).data_ptr(), _need(ln2[1], "ln2_beta", torch.float32).data_ptr()
    a.ln_eps = float(ln_eps)
    a.M, a.C, a.N3 = M, 256, N3
    a.w_static = 1 if (W1.static and W2.static and W3.static) else 0
    if out_f32 is None:
        out_f32 = torch.empty((M, N3), device=X.device, dtype=torch.float32)
    o, ldo = _rows(_need(out_f32, "out_f32", torch.float32), "out_f32")
    if o.shape != (M, N3):
        raise RuntimeError(f"transcar_b200.mlp: out_f32 has shape {tuple(o.shape)}, expected {(M, N3)}")
    a.out_f32, a.ld_out_f32 = o.data_ptr(), ldo
    keep = [X, w1, w2, w3]
    if tail is not None:
        _fill_tail(a, tail, M, X.device, keep)
    label = "mlp" if TIMELINE is None else (
        f"mlp M{M} C256 N{N3} bf16x3" + ("+ln" if ln1 is not None or ln2 is not None else "") + (f"+tail:{tail['kind']}" if tail is not None else ""))
    _lib.check(_call(label, lib.tc_mlp, C.byref(a), _stream()), "mlp")
    return out_f32


def point_embed(x, weight, bias, ln_gamma, ln_beta, logit_input, ln_eps=1e-5, want_f32=True, want_bf16=False,
                out16="bf16"):
    """ReLU(LN(Linear_{3->C}(f(x[:, :3])))); x [M, ldx>=3] fp32.  ``out16``: 'bf16' or 'split'."""
    lib = _lib.load()
    x, ldx = _rows(_need(x, "x", torch.float32), "x")
    M, Cc = x.shape[0], weight.shape[0]
    a = _lib.PointEmbedArgs()
    a.x, a.ldx, a.M, a.C, a.logit_input = x.data_ptr(), ldx, M, Cc, 1 if logit_input else 0
    a.weight = _need(weight, "weight", torch.float32).contiguous().data_ptr()
    a.bias = _need(bias, "bias", torch.float32).data_ptr()
    a.ln_gamma = _need(ln_gamma, "ln_gamma", torch.float32).data_ptr()
    a.ln_beta = _need(ln_beta, "ln_beta", torch.float32).data_ptr()
    a.ln_eps = float(ln_eps)
    o32 = torch.empty((M, Cc), device=x.device, dtype=torch.float32) if want_f32 else None
    o16 = r16 = None
    if want_bf16:
        o16, r16, a.out16_dtype = _out16(out16, (M, Cc), x.device)
    a.out_f32, a.out_bf16 = _ptr(o32), _ptr(o16)
    _lib.check(_call(f"point_embed M{M}", lib.tc_point_embed, C.byref(a), _stream()), "point_embed")
    return o32, r16


# --------------------------------------------------------------------------------------- K4
ATTN_ALGOS = {"auto": _lib.TC_ATTN_AUTO, "tensor": _lib.TC_ATTN_TENSOR, "simt": _lib.TC_ATTN_SIMT, "sparse": _lib.TC_ATTN_SPARSE}


def _set_dropout(a, dropout):
    if dropout is not None and dropout[0] > 0:
        a.dropout_p, a.dropout_seed, a.dropout_stream = float(dropout[0]), int(dropout[1]), int(dropout[2])


def attention(q, k, v, heads, *, geom=None, key_xy=None, out_dtype=None, want_row_any=False, scale=None, out=None,
              algo="auto", dropout=None, attn_blocked=None, key_blocked=None):
    """q [B,Lq,E], k/v [B,Lk,E] (fp32, bf16 or fp16; views with a row stride are fine) -> out [B,Lq,E], row_any [B,Lq]
    or None.  ``out_dtype``: a torch dtype or ``"split"`` (SplitBf16)."""
    lib = _lib.load()
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _need(t, n)
        if t.dim() != 3:
            raise RuntimeError(f"transcar_b200.attention: `{n}` must be [B, L, E]")
    B, Lq, E = q.shape
    Lk = k.shape[1]
    D = E // heads
    out_dtype = out_dtype or q.dtype
    ret = out
    if out is None:
        if out_dtype == "split":
            out, ret, _ = _out16("split", (B, Lq, E), q.device)
        else:
            ret = out = torch.empty((B, Lq, E), device=q.device, dtype=out_dtype)
    elif isinstance(out, SplitBf16):
        out = out.t
    a = _lib.AttentionArgs()
    a.q, a.k, a.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    a.ldq, a.ldk, a.ldv = q.stride(1), k.stride(1), v.stride(1)
    a.q_batch_stride, a.k_batch_stride, a.v_batch_stride = q.stride(0), k.stride(0), v.stride(0)
    a.qkv_dtype = _DT[q.dtype]
    a.B, a.Lq, a.Lk, a.heads, a.D = B, Lq, Lk, heads, D
    a.scale = float(scale if scale is not None else 1.0 / math.sqrt(D))
    if geom is not None:
        geom = _need(geom, "geom", torch.float32).contiguous()
        key_xy = _need(key_xy, "key_xy", torch.float32).contiguous()
        assert geom.shape[-1] == 8 and geom.numel() == B * Lq * 8 and key_xy.numel() == B * Lk * 2
        a.geom, a.key_xy = geom.data_ptr(), key_xy.data_ptr()
    row_any = torch.empty((B, Lq), device=q.device, dtype=torch.uint8) if want_row_any else None
    a.out, a.ldo = out.data_ptr(), out.stride(1)
    a.out_dtype = TC_BF16X2 if isinstance(ret, SplitBf16) else _DT[out.dtype]
    a.row_any = _ptr(row_any)
    a.algo = ATTN_ALGOS[algo]
    _set_dropout(a, dropout)          # (p, seed, stream): attention-probability dropout of the training variant
    if attn_blocked is not None:      # [Lq, Lk] uint8 / bool, 1 = blocked (nn.MultiheadAttention attn_mask)
        attn_blocked = _need(attn_blocked.to(torch.uint8), "attn_blocked", torch.uint8).contiguous()
        assert attn_blocked.shape == (Lq, Lk)
        a.attn_blocked = attn_blocked.data_ptr()
    if key_blocked is not None:       # [B, Lk] (key_padding_mask)
        key_blocked = _need(key_blocked.to(torch.uint8), "key_blocked", torch.uint8).contiguous()
        assert key_blocked.shape == (B, Lk)
        a.key_blocked = key_blocked.data_ptr()
    _lib.check(_call(f"attention Lq{Lq} Lk{Lk} {algo}{' masked' if geom is not None else ''}", lib.tc_attention_fwd, C.byref(a), _stream()), "attention")
    return ret, row_any


def radar_geometry(centre, code, pc_range, r_lo, r_hi, centre_is_normalised):
    """centre [M, >=2], code [M, >=8] fp32 -> geom [M, 8]."""
    lib = _lib.load()
    centre, ldc = _rows(_need(centre, "centre", torch.float32), "centre")
    code, ldk = _rows(_need(code, "code", torch.float32), "code")
    M = centre.shape[0]
    geom = torch.empty((M, 8), device=centre.device, dtype=torch.float32)
    a = _lib.RadarGeometryArgs()
    a.centre, a.ld_centre, a.centre_is_normalised = centre.data_ptr(), ldc, 1 if centre_is_normalised else 0
    a.code, a.ld_code, a.M = code.data_ptr(), ldk, M
    for i in range(6):
        a.pc_range[i] = float(pc_range[i])
    a.r_lo, a.r_hi = float(r_lo), float(r_hi)
    a.geom = geom.data_ptr()
    _lib.check(_call("radar_geometry", lib.tc_radar_geometry, C.byref(a), _stream()), "radar_geometry")
    return geom


def radar_mask(geom, key_xy, B, Lq, Lk, want_blocked=True):
    """Materialised ``blocked [B,Lq,Lk]`` uint8 (1 = not attended) + ``row_any [B,Lq]`` - test aid."""
    lib = _lib.load()
    geom = _need(geom, "geom", torch.float32).contiguous()
    key_xy = _need(key_xy, "key_xy", torch.float32).contiguous()
    blocked = torch.empty((B, Lq, Lk), device=geom.device, dtype=torch.uint8) if want_blocked else None
    row_any = torch.empty((B, Lq), device=geom.device, dtype=torch.uint8)
    _lib.check(lib.tc_radar_mask(_ptr(geom), _ptr(key_xy), B, Lq, Lk, _ptr(blocked), _ptr(row_any), _stream()),
               "radar_mask")
    return blocked, row_any


# --------------------------------------------------------------------------------------- pointwise
def ref_update(code, ref, out=None):
    """code [M, >=5], ref [M,3] -> new ref [M,3] (T:195-203)."""
    lib = _lib.load()
    code, ldc = _rows(_need(code, "code", torch.float32), "code")
    ref = _need(ref, "ref", torch.float32).contiguous()
    M = code.shape[0]
    if out is None:
        out = torch.empty((M, 3), device=code.device, dtype=torch.float32)
    _lib.check(_call("ref_update", lib.tc_ref_update, _ptr(code), ldc, _ptr(ref), _ptr(out), M, _stream()), "ref_update")
    return out


def box_anchor_add(code, anchor, xy_col, z_col, xy_from_normalised, pc_range):
    """In place: code[:,0:2] += anchor_xy(metres), code[:,4] += anchor_z (H:596-600, :664-665, :722-723)."""
    lib = _lib.load()
    code, ldc = _rows(_need(code, "code", torch.float32), "code")
    anchor, lda = _rows(_need(anchor, "anchor", torch.float32), "anchor")
    pc = (C.c_float * 6)(*[float(x) for x in pc_range])
    _lib.check(_call("box_anchor_add", lib.tc_box_anchor_add, _ptr(code), ldc, _ptr(anchor), lda, xy_col, z_col,
                     1 if xy_from_normalised else 0, C.byref(pc), code.shape[0], _stream()),
               "box_anchor_add")
    return code


def cast_bf16(src):
    lib = _lib.load()
    s, lds = _rows(_need(src, "src", torch.float32), "src")
    dst = torch.empty(s.shape, device=s.device, dtype=torch.bfloat16)
    _lib.check(lib.tc_cast_bf16(_ptr(s), lds, _ptr(dst), dst.stride(0) if dst.shape[0] > 1 else dst.shape[1],
                                s.shape[0], s.shape[1], _stream()), "cast_bf16")
    return dst.view(src.shape)


def cast_split(src):
    """fp32 [rows, cols] -> :class:`SplitBf16` (hi | lo)."""
    lib = _lib.load()
    s, lds = _rows(_need(src, "src", torch.float32), "src")
    dst = torch.empty((s.shape[0], 2 * s.shape[1]), device=s.device, dtype=torch.bfloat16)
    _lib.check(lib.tc_cast_split(_ptr(s), lds, _ptr(dst), dst.stride(0) if dst.shape[0] > 1 else dst.shape[1],
                                 s.shape[0], s.shape[1], _stream()), "cast_split")
    return SplitBf16(dst.view(*src.shape[:-1], 2 * src.shape[-1]))


def decode(cls, code, max_num, post_center_range, records=False):
    """cls [B,Q,classes], code [B,Q,10] fp32 -> boxes [B,max_num,9], scores, labels (int32), keep (uint8); with
    ``records=True`` instead ONE tensor [B,max_num,12] = (box, score, label, keep) - the result-gather record."""
    lib = _lib.load()
    cls = _need(cls, "cls", torch.float32).contiguous()
    code = _need(code, "code", torch.float32).contiguous()
    B, Q, classes = cls.shape
    if code.shape != (B, Q, 10):
        raise RuntimeError(f"transcar_b200.decode: box codes must be [B, Q, 10] (got {tuple(code.shape)})")
    dev = cls.device
    if records:
        rec = torch.empty((B, max_num, 12), device=dev, dtype=torch.float32)
        a = _lib.DecodeArgs()
        a.cls, a.code, a.B, a.Q, a.classes, a.max_num = cls.data_ptr(), code.data_ptr(), B, Q, classes, max_num
        for i in range(6):
            a.post_center_range[i] = float(post_center_range[i])
        a.records = rec.data_ptr()
        _lib.check(_call("decode", lib.tc_decode, C.byref(a), _stream()), "decode")
        return rec
    boxes = torch.empty((B, max_num, 9), device=dev, dtype=torch.float32)
    scores = torch.empty((B, max_num), device=dev, dtype=torch.float32)
    labels = torch.empty((B, max_num), device=dev, dtype=torch.int32)
    keep = torch.empty((B, max_num), device=dev, dtype=torch.uint8)
    a = _lib.DecodeArgs()
    a.cls, a.code, a.B, a.Q, a.classes, a.max_num = cls.data_ptr(), code.data_ptr(), B, Q, classes, max_num
    for i in range(6):
        a.post_center_range[i] = float(post_center_range[i])
    a.boxes, a.scores, a.labels, a.keep = boxes.data_ptr(), scores.data_ptr(), labels.data_ptr(), keep.data_ptr()
    _lib.check(_call("decode", lib.tc_decode, C.byref(a), _stream()), "decode")
    return boxes, scores, labels, keep


# --------------------------------------------------------------------------------------- backward building blocks
def transpose(src, out_dtype=None, out=None):
    """[R,C] -> [C,R] (fp32 / bf16 on either side)."""
    lib = _lib.load()
    s2, lds = _rows(_need(src, "src"), "src")
    R, Cc = s2.shape
    if out is None:
        out = torch.empty((Cc, R), device=s2.device, dtype=out_dtype or s2.dtype)
    o2, ldo = _rows(_need(out, "out"), "out")
    _lib.check(_call("transpose", lib.tc_transpose, _ptr(s2), _DT[s2.dtype], lds, _ptr(o2), _DT[o2.dtype], ldo, R, Cc, _stream()),
               "transpose")
    return out


def transpose_split(src, pad_to=64):
    """fp32 [R,C] -> :class:`SplitBf16` of the transpose, logical [C, Rp] with Rp = R rounded up to ``pad_to`` (zero padded):
    an operand of a bf16x3 GEMM that reduces over R."""
    lib = _lib.load()
    s2, lds = _rows(_need(src, "src", torch.float32), "src")
    R, Cc = s2.shape
    Rp = (R + pad_to - 1) // pad_to * pad_to
    out = torch.empty((Cc, 2 * Rp), device=s2.device, dtype=torch.bfloat16)
    _lib.check(_call("transpose_split", lib.tc_transpose, _ptr(s2), TC_F32, lds, _ptr(out), TC_BF16X2, 2 * Rp, R, Cc, _stream()),
               "transpose_split")
    return SplitBf16(out)


def colsum_(x, out):
    """out[n] += sum_m x[m,n]  (out: fp32 [N], accumulated in place)."""
    lib = _lib.load()
    x2, ldx = _rows(_need(x, "x"), "x")
    _need(out, "out", torch.float32)
    _lib.check(_call("colsum", lib.tc_colsum, _ptr(x2), _DT[x2.dtype], ldx, x2.shape[0], x2.shape[1], _ptr(out), _stream()), "colsum")
    return out


def layernorm_fwd(x, gamma, beta, eps=1e-5, relu=False, want_bf16=False, save_stats=True):
    """y = LN(x)*gamma+beta (optional ReLU); returns (y_f32, y_bf16 or None, mean, rstd)."""
    lib = _lib.load()
    x2, ldx = _rows(_need(x, "x", torch.float32), "x")
    M, N = x2.shape
    y = torch.empty((M, N), device=x2.device, dtype=torch.float32)
    y16 = torch.empty((M, N), device=x2.device, dtype=torch.bfloat16) if want_bf16 else None
    mean = torch.empty((M,), device=x2.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty((M,), device=x2.device, dtype=torch.float32) if save_stats else None
    a = _lib.LayerNormArgs()
    a.x, a.ldx, a.M, a.N = x2.data_ptr(), ldx, M, N
    a.gamma, a.beta = _need(gamma, "gamma", torch.float32).data_ptr(), _need(beta, "beta", torch.float32).data_ptr()
    a.eps, a.relu = float(eps), 1 if relu else 0
    a.y_f32, a.y_bf16, a.ldy = y.data_ptr(), _ptr(y16), N
    a.mean, a.rstd = _ptr(mean), _ptr(rstd)
    _lib.check(_call("layernorm_fwd", lib.tc_layernorm_fwd, C.byref(a), _stream()), "layernorm_fwd")
    return y, y16, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma=None, dbeta=None, add=None):
    """dx of LayerNorm (+ add); dgamma / dbeta (fp32 [N]) are accumulated in place."""
    lib = _lib.load()
    dy2, ld_dy = _rows(_need(dy, "dy", torch.float32), "dy")
    x2, ldx = _rows(_need(x, "x", torch.float32), "x")
    M, N = x2.shape
    dx = torch.empty((M, N), device=x2.device, dtype=torch.float32)
    a = _lib.LayerNormBwdArgs()
    a.dy, a.ld_dy, a.x, a.ldx = dy2.data_ptr(), ld_dy, x2.data_ptr(), ldx
    a.mean, a.rstd, a.gamma = mean.data_ptr(), rstd.data_ptr(), _need(gamma, "gamma", torch.float32).data_ptr()
    a.M, a.N = M, N
    if add is not None:
        ad, ld_add = _rows(_need(add, "add", torch.float32), "add")
        a.add, a.ld_add = ad.data_ptr(), ld_add
    a.dx, a.ld_dx = dx.data_ptr(), N
    a.dgamma, a.dbeta = _ptr(dgamma), _ptr(dbeta)
    _lib.check(_call("layernorm_bwd", lib.tc_layernorm_bwd, C.byref(a), _stream()), "layernorm_bwd")
    return dx


def mask_grad(dy, y=None, gate=None):
    """dy * (y > 0) * gate[row]  -> new fp32 tensor."""
    lib = _lib.load()
    dy2, ld_dy = _rows(_need(dy, "dy", torch.float32), "dy")
    M, N = dy2.shape
    dz = torch.empty((M, N), device=dy2.device, dtype=torch.float32)
    y_ptr, ld_y = None, 0
    if y is not None:
        y2, ld_y = _rows(_need(y, "y", torch.float32), "y")
        y_ptr = _ptr(y2)
    _lib.check(_call("mask_grad", lib.tc_mask_grad, _ptr(dy2), ld_dy, y_ptr, ld_y, _ptr(gate), _ptr(dz), N, M, N, _stream()),
               "mask_grad")
    return dz


def attention_sparse_bwd(q, k, v, dout, heads, geom, key_xy, dk, dv, scale=None, dropout=None):
    """Backward of the masked radar attention core (fp32).  q/dout [B,Lq,E]; k/v [B,Lk,E] views; dk/dv are fp32 views of the
    same shape as k/v and are ACCUMULATED.  Returns dq [B,Lq,E]."""
    lib = _lib.load()
    B, Lq, E = q.shape
    Lk = k.shape[1]
    D = E // heads
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (dout, "dout"), (dk, "dk"), (dv, "dv")):
        _need(t, n, torch.float32)
    dout = dout.contiguous()
    dq = torch.empty((B, Lq, E), device=q.device, dtype=torch.float32)
    a = _lib.AttentionBwdArgs()
    a.q, a.k, a.v, a.dout = q.data_ptr(), k.data_ptr(), v.data_ptr(), dout.data_ptr()
    a.ldq, a.ldk, a.ldv, a.ld_dout = q.stride(1), k.stride(1), v.stride(1), dout.stride(1)
    a.q_batch_stride, a.k_batch_stride, a.v_batch_stride = q.stride(0), k.stride(0), v.stride(0)
    a.B, a.Lq, a.Lk, a.heads, a.D = B, Lq, Lk, heads, D
    a.scale = float(scale if scale is not None else 1.0 / math.sqrt(D))
    a.geom, a.key_xy = _need(geom, "geom", torch.float32).contiguous().data_ptr(), _need(key_xy, "key_xy", torch.float32).contiguous().data_ptr()
    a.dq, a.ld_dq = dq.data_ptr(), E
    a.dk, a.dv = dk.data_ptr(), dv.data_ptr()
    a.ld_dk, a.ld_dv, a.dk_batch_stride, a.dv_batch_stride = dk.stride(1), dv.stride(1), dk.stride(0), dv.stride(0)
    _set_dropout(a, dropout)
    _lib.check(_call("attention_sparse_bwd", lib.tc_attention_sparse_bwd, C.byref(a), _stream()), "attention_sparse_bwd")
    return dq


def sample_bwd(feats, ref, lidar2img, attn_logits, pc_range, img_w, img_h, dout, want_feat_grad=False, d_feats=None,
               want_logit_grad=True, want_ref_grad=False):
    """Backward of :func:`sample_fwd` (``tc_sample_bwd``).  ``dout`` [B,Q,C] fp32.  Returns
    ``(d_feats or None, d_logits [B,Q,N*L] or None, d_ref [B,Q,3] or None)``; ``d_feats`` are fp32 channels-last maps that
    are ACCUMULATED into (pass existing ones through ``d_feats`` or let ``want_feat_grad`` allocate zeroed ones)."""
    lib = _lib.load()
    a = _lib.SampleBwdArgs()
    B, N, Cc = feats[0].shape[:3]
    Q = ref.shape[1]
    if len(feats) != 4:
        raise RuntimeError("transcar_b200.sample_bwd: exactly 4 feature levels are supported")
    if want_feat_grad and d_feats is None:
        d_feats = [torch.zeros((B, N, f.shape[3], f.shape[4], Cc), device=f.device, dtype=torch.float32).permute(0, 1, 4, 2, 3)
                   for f in feats]
    for l, f in enumerate(feats):
        _need(f, f"feats[{l}]", last_contig=False)
        if not is_channels_last_5d(f):
            raise RuntimeError(f"transcar_b200.sample_bwd: feats[{l}] is not channels-last")
        a.feat[l] = f.data_ptr()
        a.H[l], a.W[l] = f.shape[3], f.shape[4]
        if d_feats is not None:
            g = _need(d_feats[l], f"d_feats[{l}]", torch.float32, last_contig=False)
            if not is_channels_last_5d(g) or g.shape != f.shape:
                raise RuntimeError(f"transcar_b200.sample_bwd: d_feats[{l}] must be a channels-last fp32 map shaped like feats[{l}]")
            a.d_feat[l] = g.data_ptr()
    ref = _need(ref, "ref", torch.float32).contiguous()
    lidar2img = _need(lidar2img, "lidar2img", torch.float32).contiguous()
    attn_logits = _need(attn_logits, "attn_logits", torch.float32).contiguous()
    dout = _need(dout, "dout", torch.float32).contiguous()
    assert ref.shape == (B, Q, 3) and attn_logits.shape == (B, Q, N * 4) and dout.shape == (B, Q, Cc)
    a.num_levels, a.B, a.N, a.Q, a.C = 4, B, N, Q, Cc
    a.feat_dtype = _DT[feats[0].dtype]
    a.ref, a.lidar2img, a.attn_logits, a.dout = ref.data_ptr(), lidar2img.data_ptr(), attn_logits.data_ptr(), dout.data_ptr()
    for i in range(6):
        a.pc_range[i] = float(pc_range[i])
    a.img_w, a.img_h = float(img_w), float(img_h)
    d_logits = torch.empty((B, Q, N * 4), device=ref.device, dtype=torch.float32) if want_logit_grad else None
    d_ref = torch.empty((B, Q, 3), device=ref.device, dtype=torch.float32) if want_ref_grad else None
    a.d_logits, a.d_ref = _ptr(d_logits), _ptr(d_ref)
    _lib.check(_call("sample_bwd", lib.tc_sample_bwd, C.byref(a), _stream()), "sample_bwd")
    return d_feats, d_logits, d_ref


def attention_dense_bwd(q, k, v, o, dout, heads, scale=None, dropout=None):
    """Backward of the mask-free attention core (fp32): q/o/dout [B,Lq,E], k/v [B,Lk,E] (strided views are fine) ->
    (dq [B,Lq,E], dk [B,Lk,E], dv [B,Lk,E]) contiguous."""
    lib = _lib.load()
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (o, "o"), (dout, "dout")):
        _need(t, n, torch.float32)
    B, Lq, E = q.shape
    Lk = k.shape[1]
    D = E // heads
    dq = torch.empty((B, Lq, E), device=q.device, dtype=torch.float32)
    dk = torch.empty((B, Lk, E), device=q.device, dtype=torch.float32)
    dv = torch.empty((B, Lk, E), device=q.device, dtype=torch.float32)
    ws = torch.empty((2, B, heads, Lq), device=q.device, dtype=torch.float32)
    a = _lib.AttentionDenseBwdArgs()
    a.q, a.k, a.v, a.o, a.dout = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), dout.data_ptr()
    a.ldq, a.ldk, a.ldv, a.ldo, a.ld_dout = q.stride(1), k.stride(1), v.stride(1), o.stride(1), dout.stride(1)
    a.q_batch_stride, a.k_batch_stride, a.v_batch_stride = q.stride(0), k.stride(0), v.stride(0)
    a.o_batch_stride, a.dout_batch_stride = o.stride(0), dout.stride(0)
    a.B, a.Lq, a.Lk, a.heads, a.D = B, Lq, Lk, heads, D
    a.scale = float(scale if scale is not None else 1.0 / math.sqrt(D))
    a.dq, a.dk, a.dv, a.workspace = dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), ws.data_ptr()
    _set_dropout(a, dropout)
    _lib.check(_call("attention_dense_bwd", lib.tc_attention_dense_bwd, C.byref(a), _stream()), "attention_dense_bwd")
    return dq, dk, dv


def _pointwise(grad, x, mode, what):
    lib = _lib.load()
    x = _need(x, "x", torch.float32).contiguous()
    if grad is not None:
        grad = _need(grad, "grad", torch.float32).contiguous()
    out = torch.empty_like(x)
    _lib.check(_call(what, lib.tc_pointwise, _ptr(grad), _ptr(x), _ptr(out), x.numel(), mode, _stream()), what)
    return out


def logit_bwd(grad, x):
    """d/dx of inverse_sigmoid (T:17-32) applied to an upstream gradient."""
    return _pointwise(grad, x, 0, "logit_bwd")


def sigmoid_bwd(grad, y):
    """grad * y * (1 - y) for y = sigmoid(.)."""
    return _pointwise(grad, y, 1, "sigmoid_bwd")


def logit(x):
    """inverse_sigmoid (T:17-32)."""
    return _pointwise(None, x, 2, "logit")


def sigmoid(x):
    return _pointwise(None, x, 3, "sigmoid")


def dropout(x, p, seed, stream, residual=None, out=None):
    """out = (residual or 0) + keep * x / (1 - p) with the regenerable Philox mask of (seed, stream); x fp32 [M,N] contiguous."""
    lib = _lib.load()
    x = _need(x, "x", torch.float32)
    assert x.is_contiguous() and x.dim() == 2
    if residual is not None:
        residual = _need(residual, "residual", torch.float32)
        assert residual.is_contiguous() and residual.shape == x.shape
    out = torch.empty_like(x) if out is None else out
    _lib.check(_call("dropout", lib.tc_dropout, _ptr(x), _ptr(residual), _ptr(out), x.shape[0], x.shape[1], float(p), int(seed),
                     int(stream), _stream()), "dropout")
    return out


def add_rows(a, b, period=None, out=None):
    """out[m,:] = a[m,:] + b[m % period,:] for contiguous fp32 [M,N] / [period,N] (``out`` may be ``a``)."""
    lib = _lib.load()
    a = _need(a, "a", torch.float32)
    b = _need(b, "b", torch.float32)
    assert a.is_contiguous() and b.is_contiguous() and a.dim() == 2 and b.shape[1] == a.shape[1]
    period = period or b.shape[0]
    out = torch.empty_like(a) if out is None else out
    _lib.check(_call("add_rows", lib.tc_add_rows, _ptr(a), _ptr(b), _ptr(out), a.shape[0], a.shape[1], period, _stream()), "add_rows")
    return out


def period_sum_(x, out):
    """out[r,:] += sum_b x[b*period + r,:]; x [B*period, N], out [period, N] fp32 contiguous."""
    lib = _lib.load()
    x, out = _need(x, "x", torch.float32), _need(out, "out", torch.float32)
    assert x.is_contiguous() and out.is_contiguous() and x.shape[1] == out.shape[1] and x.shape[0] % out.shape[0] == 0
    _lib.check(_call("period_sum", lib.tc_period_sum, _ptr(x), _ptr(out), x.shape[0] // out.shape[0], out.shape[0], out.shape[1],
                     _stream()), "period_sum")
    return out
