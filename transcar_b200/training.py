"""Training variant of the TransCAR fusion head: forward + backward of everything that trains.

The reference recipe (``tools/train.py:238-252``) freezes the backbone, neck, the DETR3D transformer, ``cls_branches``,
``reg_branches`` and ``query_embedding``; gradients exist only for the radar head (reference
``projects/mmdet3d_plugin/models/dense_heads/detr3d_head.py``, H below):

    radar_position_encoder / radar_feat_encoder (H:531-536), rf_multihead_attn{,2,3} (H:578/649/707),
    rf_linear{1,2}{,_2,_3} + rf_norm{2,3}{,_2,_3} (H:583-586), final_cls{,2,3} / final_reg{,2,3} (H:592-593)

so the decoder runs through the inference engine (no grad) and this module runs the radar head with every intermediate
kept, then walks it backwards.  All math is library kernels: forward GEMMs ``tc_linear``, LayerNorm ``tc_layernorm_fwd``,
masked attention ``tc_attention_fwd`` (sparse); backward GEMMs are ``tc_linear`` on transposed operands
(``dX = dY W`` with the running gradient fused in as the residual, ``dW = dY^T X``), plus ``tc_colsum``,
``tc_layernorm_bwd``, ``tc_mask_grad`` and ``tc_attention_sparse_bwd``.  Masks are not differentiable; the only cross-layer
gradient paths are the residual stream and ``reg_l[:, (0,1,4)] += reg_{l-1}[:, (0,1,4)]`` (H:664-665, H:722-723), whose
backward is one ``tc_box_anchor_add`` on the gradients.

GEMM arithmetic: ``tensor_cores=True`` (default whenever the head runs a tensor-core precision mode) evaluates every
forward, dgrad and wgrad GEMM whose reduction length allows it on tcgen05 in bf16x3 (split-bf16 operands, fp32
accumulation: ~1e-5 of fp32, the reference trains this head in fp32); the operands of the backward GEMMs - dY^T, X^T, W^T -
are written directly as split-bf16 matrices by ``tc_transpose`` (the wgrad reduction over the B*Q rows is zero-padded to a
multiple of 64).  ``tensor_cores=False`` keeps everything on the exact fp32 CUDA-core path (parity mode).

Data parallel: every rank runs its own samples; ``sharding.GradBucket`` all-reduces the flat gradient buffer once per step.
"""
from __future__ import annotations

import torch

from . import ops
from .engine import RADIUS_CLAMP
from .radar_tokens import NUM_RADAR_FEATS

_SUF = ("", "_2", "_3")
_NUM = ("", "2", "3")


def trainable_names(state_dict_keys):
    """Parameter names that train under the reference recipe (and have a gradient path: ``rf_norm1*``,
    ``attention_weights{2,3}`` and ``output_proj{2,3}`` are never used in forward, H:135,191-195)."""
    pre = ("radar_position_encoder.", "radar_feat_encoder.", "rf_multihead_attn", "rf_linear", "rf_norm2", "rf_norm3",
           "final_cls", "final_reg")
    return [k for k in state_dict_keys if k.startswith(pre)]


def decoder_trainable_names(state_dict_keys):
    """Parameters with a gradient path when the DETR3D decoder is NOT frozen: the six decoder layers, the reference-point
    Linear and the query embedding.  ``cls_branches`` / ``reg_branches`` stay without gradient in TransCAR: their outputs
    reach the loss only through detached reference points (T:203) and the thresholded radar mask (H:543-571)."""
    return [k for k in state_dict_keys if k.startswith("transformer.") or k == "query_embedding.weight"]


class _TapeOps:
    """Taped Linear / LayerNorm building blocks shared by the radar-head and the decoder trainers: parameters by reference
    key name in ``self.p``, gradients accumulated into views of one flat buffer (``self.g``)."""

    def _setup(self, params, names, num_heads, pc_range, tensor_cores):
        self.p = params
        self.tc = tensor_cores
        self._wcache = {}                   # split-bf16 copies of the weights, valid for one forward / backward
        self.heads = num_heads
        self.pc_range = [float(v) for v in pc_range]
        self.names = names
        dev = next(iter(params.values())).device
        n = sum(params[k].numel() for k in self.names)
        self.flat_grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.g = {}
        off = 0
        for k in self.names:
            t = params[k]
            self.g[k] = self.flat_grad[off: off + t.numel()].view_as(t)
            off += t.numel()
        self.ctx = None
        self.drop_p, self.drop_seed, self.drop_step = 0.0, 0, 0

    # ------------------------------------------------------------------ dropout (training mode of the reference)
    # Sites and probabilities of the reference: p = 0.1 on the attention probabilities and on the attention output of the
    # decoder self-attention (mmcv MultiheadAttention, cfg :68-72) and of rf_multihead_attn* / rf_dropout2* (H:128-145,
    # 578-581), on Detr3DCrossAtten's output (T:378), after the FFN activation and on the FFN output (mmcv FFN ffn_dropout /
    # rf_dropout*, rf_dropout3*: H:583-585).  Masks are regenerated from (seed, stream) in the backward pass; stream =
    # step * 1024 + layer * 8 + kind with layers 0-5 = decoder, 6-8 = radar, kind: 0 attention probabilities,
    # 1 attention output, 2 cross-attention output, 3 FFN hidden, 4 FFN output.
    KIND_PROBS, KIND_ATTN_OUT, KIND_CROSS_OUT, KIND_FFN_HIDDEN, KIND_FFN_OUT = 0, 1, 2, 3, 4

    def set_dropout(self, p=0.0, seed=0, step=0):
        self.drop_p, self.drop_seed, self.drop_step = float(p), int(seed), int(step)

    def _stream(self, layer, kind):
        return self.drop_step * 1024 + layer * 8 + kind

    def _drop_arg(self, layer, kind):
        """(p, seed, stream) for the attention kernels, or None when dropout is off."""
        return (self.drop_p, self.drop_seed, self._stream(layer, kind)) if self.drop_p > 0 else None

    # ------------------------------------------------------------------ GEMM plumbing
    def _wsplit(self, wkey, w_rows, W, transposed):
        """Split-bf16 copy of a weight (or of its transpose) for this step; the parameters change between steps."""
        key = (wkey, w_rows, transposed)
        t = self._wcache.get(key)
        if t is None:
            t = self._wcache[key] = ops.transpose_split(W) if transposed else ops.cast_split(W)
        return t

    # ------------------------------------------------------------------ differentiable pieces (forward records a tape)
    def _linear(self, tape, x, wkey, bkey, relu=False, residual=None, gate=None, w_rows=None, residual2=None, drop=None):
        """y = [relu]( gate * (x W^T + b) + residual + residual2 ).  ``w_rows`` selects a row range of a packed
        weight/bias.  The residual terms pass their gradient through unchanged (the caller routes it).
        ``drop = (layer, kind)``: dropout on the Linear's (activated, gated) output BEFORE the residuals are added."""
        W, b = self.p[wkey], self.p[bkey]
        if w_rows is not None:
            W, b = W[w_rows[0]:w_rows[1]], b[w_rows[0]:w_rows[1]]
        if drop is not None and self.drop_p > 0:
            stream = self._stream(*drop)
            t = self._linear([], x, wkey, bkey, relu=relu, gate=gate, w_rows=w_rows)
            y = ops.dropout(t, self.drop_p, self.drop_seed, stream, residual=residual)
            if residual2 is not None:
                y = ops.add_rows(y, residual2.contiguous(), y.shape[0], out=y)
            tape.append(("linear", x, wkey, bkey, w_rows, t if relu else None, gate, stream))
            return y
        if self.tc and x.shape[1] % 64 == 0:
            y, _ = ops.linear(ops.cast_split(x), self._wsplit(wkey, w_rows, W, False), b, relu=relu, residual=residual,
                              residual2=residual2, row_gate=gate)
        else:               # K = 3 / 36 (raw radar fields, reference points): exact fp32 CUDA-core path
            y, _ = ops.linear(x, W, b, relu=relu, residual=residual, residual2=residual2, row_gate=gate)
        tape.append(("linear", x, wkey, bkey, w_rows, y if relu else None, gate, None))
        return y

    def _linear_bwd(self, rec, dy, dx_accum=None, need_dx=True):
        _, x, wkey, bkey, w_rows, y_relu, gate, stream = rec
        if stream is not None:               # same mask as the forward, regenerated
            dy = ops.dropout(dy.contiguous(), self.drop_p, self.drop_seed, stream)
        if y_relu is not None or gate is not None:
            dy = ops.mask_grad(dy, y=y_relu, gate=gate)
        W, gW, gb = self.p[wkey], self.g[wkey], self.g[bkey]
        if w_rows is not None:
            W, gW, gb = W[w_rows[0]:w_rows[1]], gW[w_rows[0]:w_rows[1]], gb[w_rows[0]:w_rows[1]]
        ops.colsum_(dy, gb)
        if self.tc:         # dW = dY^T X on tcgen05: both operands transposed straight into split-bf16, M padded to 64
            ops.linear(ops.transpose_split(dy), ops.transpose_split(x), None, out_f32=gW)
        else:
            dyT, xT = ops.transpose(dy), ops.transpose(x)                  # [N,M], [K,M]
            ops.linear(dyT, xT, None, out_f32=gW)                          # dW = dY^T X
        if not need_dx:
            return None
        if self.tc and dy.shape[1] % 64 == 0:
            dx, _ = ops.linear(ops.cast_split(dy), self._wsplit(wkey, w_rows, W, True), None, residual=dx_accum)
        else:               # N = 10 (box / class logits): exact fp32 CUDA-core path
            WT = ops.transpose(W)                                          # [K,N]
            dx, _ = ops.linear(dy, WT, None, residual=dx_accum)            # dX = dY W (+ running gradient)
        return dx

    def _ln(self, tape, x, key, relu=False):
        y, _, mean, rstd = ops.layernorm_fwd(x, self.p[key + ".weight"], self.p[key + ".bias"], relu=relu)
        tape.append(("ln", x, key, mean, rstd, y if relu else None))
        return y

    def _ln_bwd(self, rec, dy, add=None):
        _, x, key, mean, rstd, y_relu = rec
        if y_relu is not None:
            dy = ops.mask_grad(dy, y=y_relu)
        return ops.layernorm_bwd(dy, x, mean, rstd, self.p[key + ".weight"], dgamma=self.g[key + ".weight"],
                                 dbeta=self.g[key + ".bias"], add=add)


class RadarHeadTrainer(_TapeOps):
    """Forward / backward of the radar fusion head over batch-major ``[B*Q, C]`` activations."""

    def __init__(self, params, num_heads=8, pc_range=(-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), tensor_cores=True):
        """``params``: dict name -> fp32 CUDA tensor (the live parameters, reference key names)."""
        self._setup(params, trainable_names(params.keys()), num_heads, pc_range, tensor_cores)

    # ------------------------------------------------------------------ forward
    def forward(self, x0, ref, code, tokens, key_xy, B):
        """x0 [B*Q,C] decoder output, ref [B*Q,3] refined reference points, code [B*Q,10] = reg_branches[5](x0),
        tokens [B,R,36], key_xy [B,R,2].  Returns (cls [3,B,Q,10], reg [3,B,Q,10])."""
        M, C = x0.shape
        Q = M // B
        R = tokens.shape[1]
        dev = x0.device
        self._wcache = {}
        ctx = dict(B=B, Q=Q, R=R, C=C, enc=[], layers=[])
        # ---- radar encoders (H:531-536)
        t2 = tokens.reshape(B * R, NUM_RADAR_FEATS)
        enc = ctx["enc"]
        xyz = t2[:, :3]
        pe = self._ln(enc, self._linear(enc, xyz, "radar_position_encoder.0.weight", "radar_position_encoder.0.bias"),
                      "radar_position_encoder.1", relu=True)
        pos = self._ln(enc, self._linear(enc, pe, "radar_position_encoder.3.weight", "radar_position_encoder.3.bias"),
                       "radar_position_encoder.4", relu=True)
        n_pos = len(enc)
        f = self._linear(enc, t2, "radar_feat_encoder.0.weight", "radar_feat_encoder.0.bias", relu=True)
        f = self._linear(enc, f, "radar_feat_encoder.2.weight", "radar_feat_encoder.2.bias", relu=True)
        self._linear(enc, f, "radar_feat_encoder.4.weight", "radar_feat_encoder.4.bias", relu=True)   # taped: ReLU mask
        # kvfeat = pos + relu(feat) (H:536): the same small GEMM once more with the sum fused as its post-add epilogue
        if self.tc:
            kvfeat_sum, _ = ops.linear(ops.cast_split(f), self._wsplit("radar_feat_encoder.4.weight", None,
                                                                       self.p["radar_feat_encoder.4.weight"], False),
                                       self.p["radar_feat_encoder.4.bias"], relu=True, post_add=pos)
        else:
            kvfeat_sum, _ = ops.linear(f, self.p["radar_feat_encoder.4.weight"], self.p["radar_feat_encoder.4.bias"], relu=True,
                                       post_add=pos)
        ctx["n_pos"] = n_pos
        ctx["kvfeat"] = kvfeat_sum
        cls_all = torch.empty((3, B, Q, 10), device=dev, dtype=torch.float32)
        reg_all = torch.empty((3, B, Q, 10), device=dev, dtype=torch.float32)
        x, anchor, centre_norm = x0, ref, True
        for li in range(3):
            s, m = _SUF[li], _NUM[li]
            tape = []
            mha = "rf_multihead_attn" + m
            lo, hi = RADIUS_CLAMP[li]
            geom = ops.radar_geometry(anchor, code, self.pc_range, lo, hi, centre_is_normalised=centre_norm)
            q = self._linear(tape, x, mha + ".in_proj_weight", mha + ".in_proj_bias", w_rows=(0, C))
            kv = self._linear(tape, kvfeat_sum, mha + ".in_proj_weight", mha + ".in_proj_bias", w_rows=(C, 3 * C))
            kv3 = kv.view(B, R, 2 * C)
            att, row_any = ops.attention(q.view(B, Q, C), kv3[:, :, :C], kv3[:, :, C:], self.heads, geom=geom, key_xy=key_xy,
                                         want_row_any=True, algo="sparse", dropout=self._drop_arg(6 + li, self.KIND_PROBS))
            gate = row_any.view(M)
            z2 = self._linear(tape, att.view(M, C), mha + ".out_proj.weight", mha + ".out_proj.bias", residual=x, gate=gate,
                              drop=(6 + li, self.KIND_ATTN_OUT))
            x2 = self._ln(tape, z2, "rf_norm2" + s)
            h = self._linear(tape, x2, "rf_linear1" + s + ".weight", "rf_linear1" + s + ".bias", relu=True,
                             drop=(6 + li, self.KIND_FFN_HIDDEN))
            z3 = self._linear(tape, h, "rf_linear2" + s + ".weight", "rf_linear2" + s + ".bias", residual=x2,
                              drop=(6 + li, self.KIND_FFN_OUT))
            x3 = self._ln(tape, z3, "rf_norm3" + s)
            n_trunk = len(tape)
            c = self._ln(tape, self._linear(tape, x3, f"final_cls{m}.0.weight", f"final_cls{m}.0.bias"), f"final_cls{m}.1", relu=True)
            c = self._ln(tape, self._linear(tape, c, f"final_cls{m}.3.weight", f"final_cls{m}.3.bias"), f"final_cls{m}.4", relu=True)
            cls = self._linear(tape, c, f"final_cls{m}.6.weight", f"final_cls{m}.6.bias")
            n_cls = len(tape)
            g = self._linear(tape, x3, f"final_reg{m}.0.weight", f"final_reg{m}.0.bias", relu=True)
            g = self._linear(tape, g, f"final_reg{m}.2.weight", f"final_reg{m}.2.bias", relu=True)
            reg = self._linear(tape, g, f"final_reg{m}.4.weight", f"final_reg{m}.4.bias")
            if li == 0:   # H:596-600: x,y of the refined reference in metres, z left normalised (quirk Q3)
                ops.box_anchor_add(reg, anchor, 0, 2, True, self.pc_range)
            else:         # H:664-665, H:722-723
                ops.box_anchor_add(reg, anchor, 0, 4, False, self.pc_range)
            cls_all[li].view(M, 10).copy_(cls)
            reg_all[li].view(M, 10).copy_(reg)
            ctx["layers"].append(dict(tape=tape, n_trunk=n_trunk, n_cls=n_cls, q=q, kv=kv, geom=geom, gate=gate))
            x, anchor, code, centre_norm = x3, reg, reg, False
        ctx["key_xy"] = key_xy
        self.ctx = ctx
        return cls_all, reg_all

    # ------------------------------------------------------------------ backward
    def backward(self, d_cls, d_reg, need_dx0=False):
        """d_cls / d_reg: [3,B,Q,10] upstream gradients.  Accumulates into ``self.flat_grad`` (zero it per step).
        ``need_dx0``: also return the gradient w.r.t. the decoder output ``x0`` (un-frozen decoder)."""
        ctx = self.ctx
        B, Q, R, C = ctx["B"], ctx["Q"], ctx["R"], ctx["C"]
        M = B * Q
        dev = d_cls.device
        d_kvfeat = None                     # running gradient of the radar key/value features [B*R, C]
        dx_next = None                      # gradient flowing into this layer's output x3 from the next layer
        d_reg_next = None                   # gradient of the next layer's regression (anchor path)
        for li in (2, 1, 0):
            L = ctx["layers"][li]
            tape, n_trunk, n_cls = L["tape"], L["n_trunk"], L["n_cls"]
            dreg = d_reg[li].reshape(M, 10).contiguous().clone()
            if d_reg_next is not None:      # reg_{l+1}[:, (0,1,4)] += reg_l[:, (0,1,4)]  ->  dreg_l[:, (0,1,4)] += dreg_{l+1}
                ops.box_anchor_add(dreg, d_reg_next, 0, 4, False, self.pc_range)
            d_reg_next = dreg
            # regression head (3 linears), classification head (linear, LN, linear, LN, linear): both end in dx3
            dg = self._linear_bwd(tape[n_cls + 2], dreg)
            dg = self._linear_bwd(tape[n_cls + 1], dg)
            dx3 = self._linear_bwd(tape[n_cls + 0], dg, dx_accum=dx_next)
            dc = self._linear_bwd(tape[n_trunk + 4], d_cls[li].reshape(M, 10).contiguous())
            dc = self._ln_bwd(tape[n_trunk + 3], dc)
            dc = self._linear_bwd(tape[n_trunk + 2], dc)
            dc = self._ln_bwd(tape[n_trunk + 1], dc)
            dx3 = self._linear_bwd(tape[n_trunk + 0], dc, dx_accum=dx3)
            # trunk, reversed: ln3, linear2(+res x2), linear1(relu), ln2, out_proj(+res x, gate), attention, kv, q
            dz3 = self._ln_bwd(tape[6], dx3)
            dh = self._linear_bwd(tape[5], dz3)
            dx2 = self._linear_bwd(tape[4], dh, dx_accum=dz3)               # + skip connection z3 = x2 + ...
            dz2 = self._ln_bwd(tape[3], dx2)
            datt = self._linear_bwd(tape[2], dz2)                            # gate applied inside
            kv3 = L["kv"].view(B, R, 2 * C)
            dkv = torch.zeros((B, R, 2 * C), device=dev, dtype=torch.float32)
            dq = ops.attention_sparse_bwd(L["q"].view(B, Q, C), kv3[:, :, :C], kv3[:, :, C:], datt.view(B, Q, C), self.heads,
                                          L["geom"], ctx["key_xy"], dkv[:, :, :C], dkv[:, :, C:],
                                          dropout=self._drop_arg(6 + li, self.KIND_PROBS))
            d_kvfeat = self._linear_bwd(tape[1], dkv.view(B * R, 2 * C), dx_accum=d_kvfeat)
            dx_next = self._linear_bwd(tape[0], dq.view(M, C), dx_accum=dz2, need_dx=li > 0 or need_dx0)   # + skip z2 = x + ...
        # ---- radar encoders: kvfeat = pos + relu(feat)
        enc, n_pos = ctx["enc"], ctx["n_pos"]
        df = self._linear_bwd(enc[n_pos + 2], d_kvfeat)
        df = self._linear_bwd(enc[n_pos + 1], df)
        self._linear_bwd(enc[n_pos + 0], df, need_dx=False)
        dp = self._ln_bwd(enc[3], d_kvfeat)
        dp = self._linear_bwd(enc[2], dp)
        dp = self._ln_bwd(enc[1], dp)
        self._linear_bwd(enc[0], dp, need_dx=False)
        self.ctx = None
        self._wcache = {}
        self.dx0 = dx_next if need_dx0 else None
        return self.g


class DecoderTrainer(_TapeOps):
    """Forward / backward of the six DETR3D decoder layers (T:155-214 + the mmcv layer of cfg
    ``detr3d_res101_gridmask.py:65-82``) for the un-frozen recipe: everything BASELINE.json configs[4] names - the
    grid_sample scatter (``tc_sample_bwd``) and the dense self-attention backward (``tc_attention_dense_bwd``) - plus the
    Linear / LayerNorm tape shared with the radar head.  Reference points are detached between layers (T:203), so only layer 0
    sends gradient into ``transformer.reference_points`` (through the sampling grid and the position encoder); the
    refinement branches themselves run without a tape.  Dropout: see ``_TapeOps.set_dropout``."""

    def __init__(self, params, num_query, num_heads=8, num_layers=6, pc_range=(-51.2, -51.2, -5.0, 51.2, 51.2, 3.0),
                 tensor_cores=True):
        self._setup(params, decoder_trainable_names(params.keys()), num_heads, pc_range, tensor_cores)
        self.Q, self.L = num_query, num_layers

    def _reg_branch(self, l, x):
        """reg_branches[l] without a tape: its output only feeds detached reference points and the radar-mask thresholds."""
        p = self.p
        h, _ = ops.linear(x, p[f"reg_branches.{l}.0.weight"], p[f"reg_branches.{l}.0.bias"], relu=True)
        h, _ = ops.linear(h, p[f"reg_branches.{l}.2.weight"], p[f"reg_branches.{l}.2.bias"], relu=True)
        return ops.linear(h, p[f"reg_branches.{l}.4.weight"], p[f"reg_branches.{l}.4.bias"])[0]

    def forward(self, feats, l2i, img_w, img_h, B):
        """feats: 4 channels-last maps; returns (x [B*Q,C], ref [B*Q,3], code [B*Q,10]) = decoder output, refined reference
        points, ``reg_branches[5]`` output - what the radar head consumes."""
        Q, L = self.Q, self.L
        emb = self.p["query_embedding.weight"]
        C = emb.shape[1] // 2
        M = B * Q
        self._wcache = {}
        pos_q = emb[:, :C].contiguous()                                  # T:119-121
        x = emb[:, C:].contiguous().unsqueeze(0).expand(B, Q, C).reshape(M, C).contiguous()
        head = []
        r0 = self._linear(head, pos_q, "transformer.reference_points.weight", "transformer.reference_points.bias")
        ref_q = ops.sigmoid(r0)                                          # T:122-123  [Q,3]
        ref = ref_q.unsqueeze(0).expand(B, Q, 3).reshape(M, 3).contiguous()
        ctx = dict(B=B, Q=Q, C=C, head=head, ref_q=ref_q, ref0=ref, layers=[], feats=feats, l2i=l2i, img=(img_w, img_h))
        code = None
        for l in range(L):
            pre = f"transformer.decoder.layers.{l}."
            mha, ca = pre + "attentions.0.attn.", pre + "attentions.1."
            tape = []
            # ---- self-attention: q = k = x + query_pos, v = x (mmcv MultiheadAttention wrapper)
            xp = ops.add_rows(x, pos_q, Q)
            q = self._linear(tape, xp, mha + "in_proj_weight", mha + "in_proj_bias", w_rows=(0, C))
            k = self._linear(tape, xp, mha + "in_proj_weight", mha + "in_proj_bias", w_rows=(C, 2 * C))
            v = self._linear(tape, x, mha + "in_proj_weight", mha + "in_proj_bias", w_rows=(2 * C, 3 * C))
            o, _ = ops.attention(q.view(B, Q, C), k.view(B, Q, C), v.view(B, Q, C), self.heads, algo="simt",
                                 dropout=self._drop_arg(l, self.KIND_PROBS))
            z = self._linear(tape, o.view(M, C), mha + "out_proj.weight", mha + "out_proj.bias", residual=x,
                             drop=(l, self.KIND_ATTN_OUT))
            x1 = self._ln(tape, z, pre + "norms.0")
            # ---- Detr3DCrossAtten (T:302-378): residual = x1 (quirk Q1), logits from x1 + query_pos
            xp1 = ops.add_rows(x1, pos_q, Q)
            logits = self._linear(tape, xp1, ca + "attention_weights.weight", ca + "attention_weights.bias")
            s, _ = ops.sample_fwd(feats, ref.view(B, Q, 3), l2i, logits.view(B, Q, -1), self.pc_range, img_w, img_h)
            lg = ops.logit(ref)
            pe = self._ln(tape, self._linear(tape, lg, ca + "position_encoder.0.weight", ca + "position_encoder.0.bias"),
                          ca + "position_encoder.1", relu=True)
            pf = self._ln(tape, self._linear(tape, pe, ca + "position_encoder.3.weight", ca + "position_encoder.3.bias"),
                          ca + "position_encoder.4", relu=True)
            z2 = self._linear(tape, s.view(M, C), ca + "output_proj.weight", ca + "output_proj.bias", residual=x1, residual2=pf,
                              drop=(l, self.KIND_CROSS_OUT))
            x2 = self._ln(tape, z2, pre + "norms.1")
            # ---- FFN
            h = self._linear(tape, x2, pre + "ffns.0.layers.0.0.weight", pre + "ffns.0.layers.0.0.bias", relu=True,
                             drop=(l, self.KIND_FFN_HIDDEN))
            z3 = self._linear(tape, h, pre + "ffns.0.layers.1.weight", pre + "ffns.0.layers.1.bias", residual=x2,
                              drop=(l, self.KIND_FFN_OUT))
            x3 = self._ln(tape, z3, pre + "norms.2")
            ctx["layers"].append(dict(tape=tape, q=q, k=k, v=v, o=o, ref=ref, logits=logits))
            # ---- iterative refinement (T:190-203), detached
            code = self._reg_branch(l, x3)
            ref = ops.ref_update(code, ref)
            x = x3
        self.ctx = ctx
        return x, ref, code

    def backward(self, dx, want_feat_grad=False):
        """dx [B*Q,C]: gradient w.r.t. the decoder output.  Accumulates parameter gradients into ``self.flat_grad``;
        returns the feature-map gradients (4 fp32 channels-last maps) when ``want_feat_grad``."""
        ctx = self.ctx
        B, Q, C = ctx["B"], ctx["Q"], ctx["C"]
        M = B * Q
        feats, l2i, (img_w, img_h) = ctx["feats"], ctx["l2i"], ctx["img"]
        dev = dx.device
        dpos = torch.zeros((M, C), device=dev, dtype=torch.float32)      # gradient of the batch-broadcast query_pos
        d_feats = None
        d_ref0 = None
        for l in range(self.L - 1, -1, -1):
            Lc = ctx["layers"][l]
            t = Lc["tape"]
            # tape order: 0 q, 1 k, 2 v, 3 out_proj, 4 norms.0, 5 attention_weights, 6 pe.0, 7 pe.1, 8 pe.3, 9 pe.4,
            #             10 output_proj, 11 norms.1, 12 ffn1, 13 ffn2, 14 norms.2
            dz3 = self._ln_bwd(t[14], dx)
            dh = self._linear_bwd(t[13], dz3)
            dx2 = self._linear_bwd(t[12], dh, dx_accum=dz3)
            dz2 = self._ln_bwd(t[11], dx2)                                # = d x1 (residual) = d pos_feat (residual2)
            ds = self._linear_bwd(t[10], dz2)
            # position encoder (T:377): gradient reaches the reference points only in layer 0
            dpe = self._ln_bwd(t[9], dz2)
            dpe = self._linear_bwd(t[8], dpe)
            dpe = self._ln_bwd(t[7], dpe)
            d_lg = self._linear_bwd(t[6], dpe, need_dx=l == 0)
            # sampling (T:367-373, 381-422)
            d_feats, d_logits, d_ref_s = ops.sample_bwd(feats, Lc["ref"].view(B, Q, 3), l2i, Lc["logits"].view(B, Q, -1),
                                                        self.pc_range, img_w, img_h, ds.view(B, Q, C),
                                                        want_feat_grad=want_feat_grad, d_feats=d_feats, want_ref_grad=l == 0)
            dxp1 = self._linear_bwd(t[5], d_logits.view(M, -1))
            ops.add_rows(dpos, dxp1, M, out=dpos)
            dx1 = ops.add_rows(dz2, dxp1, M)
            if l == 0:
                d_ref0 = ops.add_rows(_pad4(d_ref_s.view(M, 3)), _pad4(ops.logit_bwd(d_lg, Lc["ref"])), M)[:, :3]
            # self-attention
            dz = self._ln_bwd(t[4], dx1)
            do = self._linear_bwd(t[3], dz)
            dq, dk, dv = ops.attention_dense_bwd(Lc["q"].view(B, Q, C), Lc["k"].view(B, Q, C), Lc["v"].view(B, Q, C), Lc["o"],
                                                 do.view(B, Q, C), self.heads, dropout=self._drop_arg(l, self.KIND_PROBS))
            dxp = self._linear_bwd(t[0], dq.view(M, C))
            dxp = self._linear_bwd(t[1], dk.view(M, C), dx_accum=dxp)
            ops.add_rows(dpos, dxp, M, out=dpos)
            dxv = self._linear_bwd(t[2], dv.view(M, C), dx_accum=dz)      # + residual z = x + ...
            dx = ops.add_rows(dxv, dxp, M)
        # ---- query embedding (T:119-121) and the initial reference points (T:122-123)
        g_emb = self.g["query_embedding.weight"]
        d_query = torch.zeros((Q, C), device=dev, dtype=torch.float32)
        d_posq = torch.zeros((Q, C), device=dev, dtype=torch.float32)
        ops.period_sum_(dx, d_query)
        ops.period_sum_(dpos, d_posq)
        d_refq = torch.zeros((Q, 4), device=dev, dtype=torch.float32)
        ops.period_sum_(_pad4(d_ref0), d_refq)
        d_r0 = ops.sigmoid_bwd(d_refq[:, :3].contiguous(), ctx["ref_q"])
        d_posq = self._linear_bwd(ctx["head"][0], d_r0, dx_accum=d_posq)
        g_emb[:, :C] += d_posq
        g_emb[:, C:] += d_query
        self.ctx = None
        self._wcache = {}
        return d_feats


def _pad4(t):
    """[M,3] -> contiguous [M,4] (zero column): the row kernels work on 16-byte rows."""
    out = torch.zeros((t.shape[0], 4), device=t.device, dtype=t.dtype)
    out[:, :3] = t
    return out


class _RadarHeadFunction(torch.autograd.Function):
    """Autograd bridge: forward / backward are the library-kernel passes above; the parameters enter as inputs so that
    ``loss.backward()`` deposits their gradients like for any other module."""

    @staticmethod
    def forward(ctx, trainer, x0, ref, code, tokens, key_xy, B, *params):
        ctx.trainer = trainer
        ctx.n = len(params)
        with torch.no_grad():
            cls_all, reg_all = trainer.forward(x0, ref, code, tokens, key_xy, B)
        return cls_all, reg_all

    @staticmethod
    def backward(ctx, d_cls, d_reg):
        tr = ctx.trainer
        with torch.no_grad():
            tr.flat_grad.zero_()
            grads = tr.backward(d_cls.contiguous(), d_reg.contiguous())
        return (None,) * 7 + tuple(grads[k].clone() for k in tr.names)


class _FusionHeadFunction(torch.autograd.Function):
    """Autograd bridge of the un-frozen recipe: decoder tape + radar-head tape behind one node.  Inputs: the four feature
    maps (gradient only when they require it), then the decoder parameters, then the radar-head parameters."""

    @staticmethod
    def forward(ctx, dec, radar, l2i, img_w, img_h, tokens, key_xy, B, *tensors):
        feats = list(tensors[:4])
        ctx.dec, ctx.radar = dec, radar
        ctx.feat_grad = [f.requires_grad for f in feats]
        ctx.feat_dtype = feats[0].dtype
        with torch.no_grad():
            x0, ref, code = dec.forward([f.detach() for f in feats], l2i, img_w, img_h, B)
            cls_all, reg_all = radar.forward(x0, ref, code, tokens, key_xy, B)
        return cls_all, reg_all

    @staticmethod
    def backward(ctx, d_cls, d_reg):
        dec, radar = ctx.dec, ctx.radar
        with torch.no_grad():
            radar.flat_grad.zero_()
            dec.flat_grad.zero_()
            rg = radar.backward(d_cls.contiguous(), d_reg.contiguous(), need_dx0=True)
            d_feats = dec.backward(radar.dx0, want_feat_grad=any(ctx.feat_grad))
        fg = [(g.to(ctx.feat_dtype) if (w and g is not None) else None) for w, g in zip(ctx.feat_grad, d_feats or [None] * 4)]
        return (None,) * 8 + tuple(fg) + tuple(dec.g[k].clone() for k in dec.names) + tuple(rg[k].clone() for k in radar.names)


def fusion_head_apply(dec, radar, named_params, feats, l2i, img_w, img_h, tokens, key_xy, B):
    plist = [named_params[k] for k in dec.names] + [named_params[k] for k in radar.names]
    return _FusionHeadFunction.apply(dec, radar, l2i, img_w, img_h, tokens, key_xy, B, *feats, *plist)


def radar_head_apply(trainer, named_params, x0, ref, code, tokens, key_xy, B):
    """``named_params``: dict name -> nn.Parameter (live); order follows ``trainer.names``."""
    plist = [named_params[k] for k in trainer.names]
    return _RadarHeadFunction.apply(trainer, x0, ref, code, tokens, key_xy, B, *plist)
