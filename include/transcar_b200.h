/*
 * transcar_b200 - C ABI of the B200-native (sm_100a) TransCAR fusion-decoder hot path.
 *
 * Drop-in boundary.  The reference has no native code of its own on this path:
 * its "FFI" is PyTorch's dispatcher reaching ATen/cuBLAS kernels from two Python files.  Each entry
 * point below names the reference call site(s) it replaces; paths are relative to
 * /root/reference/projects/mmdet3d_plugin/ with
 *     T = models/utils/detr3d_transformer.py      H = models/dense_heads/detr3d_head.py
 *     C = core/bbox/coders/nms_free_coder.py      U = core/bbox/util.py
 *
 * Conventions
 *   - Plain C: POD structs of raw DEVICE pointers, int32 sizes, a CUDA stream handle.  No torch types.
 *   - Ownership: the caller allocates every input, output and workspace; the library never allocates
 *     or frees device memory and keeps no pointer after a call returns.
 *   - Asynchronous: all work is enqueued on the caller's stream; no internal streams, no implicit sync.
 *     Calls are capturable into CUDA graphs.
 *   - Errors: 0 = ok, <0 = bad argument (TC_ERR_*), >0 = cudaError_t.  tc_last_error_string() describes
 *     the last failure on the calling thread.  No C++ exception crosses the boundary.
 *   - sm_100a only.  There is no CPU fallback and no other backend.
 *   - Row-major everywhere.  "fp32"/"bf16" tensors are selected by tc_dtype fields.
 */
#ifndef TRANSCAR_B200_H_
#define TRANSCAR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TC_API __attribute__((visibility("default")))
#else
#define TC_API
#endif

#define TC_ABI_VERSION 8
#define TC_MAX_LEVELS 4
#define TC_MAX_CAMS 8

typedef void* tc_stream_t;                  /* cudaStream_t */

/* TC_BF16X2 ("split bf16") stores a logical [rows, cols] matrix as bf16 [rows, 2 * cols]: columns [0, cols) hold
 * hi = bf16(x), columns [cols, 2 * cols) hold lo = bf16(x - hi) - 16 mantissa bits in two tensor-core operands.  A
 * product of two split matrices is evaluated as hi*hi + lo*hi + hi*lo on the bf16 tensor cores with fp32 accumulation
 * ("bf16x3", ~1e-5 relative: the parity-grade tensor-core mode; plain TC_BF16 is the single-pass mode, ~2e-3).
 * TC_F16 (IEEE half, 11 mantissa bits) is the operand format of the dense attention core in that mode. */
typedef enum { TC_F32 = 0, TC_BF16 = 1, TC_BF16X2 = 2, TC_F16 = 3 } tc_dtype;

enum {
  TC_OK = 0,
  TC_ERR_NULL = -1,        /* required pointer is NULL */
  TC_ERR_SHAPE = -2,       /* unsupported size */
  TC_ERR_ALIGN = -3,       /* pointer / leading dimension not aligned as documented */
  TC_ERR_DTYPE = -4,       /* unsupported dtype combination */
  TC_ERR_DEVICE = -5       /* not an sm_100 device */
};

TC_API int tc_abi_version(void);
TC_API const char* tc_last_error_string(void);
/* 0 if the current CUDA device is compute capability 10.x, TC_ERR_DEVICE otherwise. */
TC_API int tc_check_device(void);
/* Number of kernels this library has enqueued from the calling process since load (bench evidence). */
TC_API uint64_t tc_launch_count(void);
/* Profiling hook (tools/linear_trace.py): every CTA of the following tensor-core tc_linear launches writes 16 uint64
 * timestamps (globaltimer ns / clock64 cycles at its phase boundaries, SM id) into `device_buf`, one record per CTA in launch
 * order, until `records` are used up.  NULL / 0 switches it off (the default; the kernels then do no extra work). */
TC_API int tc_debug_trace(uint64_t* device_buf, int64_t records);

/* ------------------------------------------------------------------------------------------------
 * K1  fused camera sampling.   Replaces T:381-422 (feature_sampling: projection through lidar2img,
 * validity mask, 4x F.grid_sample bilinear/zeros/align_corners=False) and T:367-373 (NaN->0,
 * sigmoid(attention_weights)*mask, sum over level/point/camera) in ONE kernel: one warp per query,
 * cameras that fail the validity test are skipped, every texel is read with 128-bit loads from
 * channels-last feature maps, result written once.
 *
 *   feat[l]      [B, N, H_l, W_l, C]  channels-last, feat_dtype (fp32 or bf16), 16-byte aligned, C % 256 == 0
 *   ref          [B, Q, 3] fp32       normalised reference points in [0,1]
 *   lidar2img    [B, N, 4, 4] fp32    (reference casts the float64 matrices to fp32 at T:386)
 *   attn_logits  [B, Q, N*L] fp32     output of the attention_weights Linear, index = cam*L + level
 *   out          [B, Q, C] out_dtype  sum_{cam,level} mask * sigmoid(logit) * bilinear(feat)
 *                (fp32, bf16, or TC_BF16X2: split bf16 [B, Q, 2C], hi | lo)
 *   mask         [B, Q, N] uint8      optional (may be NULL): camera validity, bit-exact w.r.t. T:400-409
 */
/* flags: TC_SAMPLE_ALL_CAMS = cameras that fail the validity test are sampled too (only `mask` records the test) -
 * the un-masked return value of the reference's free function feature_sampling (T:381-422); the fused hot path
 * (Detr3DCrossAtten) multiplies by the mask and therefore never sets it. */
#define TC_SAMPLE_ALL_CAMS 1
/* TC_SAMPLE_WEIGHTS_GIVEN = attn_logits already holds the per-(camera, level) weights (no sigmoid is applied): how
 * num_points > 1 configurations are served - the reference samples ONE point and broadcasts it against num_points weights
 * (T:346-373), i.e. the effective weight is sum_p sigmoid(logit[cam, p, level]). */
#define TC_SAMPLE_WEIGHTS_GIVEN 2
typedef struct {
  const void* feat[TC_MAX_LEVELS];
  int32_t H[TC_MAX_LEVELS];
  int32_t W[TC_MAX_LEVELS];
  int32_t num_levels;
  int32_t B, N, Q, C;
  int32_t feat_dtype;
  int32_t out_dtype;
  const float* ref;
  const float* lidar2img;
  const float* attn_logits;
  float pc_range[6];
  float img_w, img_h;          /* img_metas[0]['img_shape'][0][1], [0][0] (quirk Q2) */
  void* out;
  uint8_t* mask;
  int32_t flags;               /* TC_SAMPLE_* bits */
} tc_sample_args;
TC_API int tc_sample_fwd(const tc_sample_args* a, tc_stream_t stream);

/* NCHW fp32 -> channels-last (fp32 or bf16) layout conversion for callers whose backbone is not
 * channels-last.  src [P, C, HW] fp32 -> dst [P, HW, C].  */
TC_API int tc_nchw_to_nhwc(const float* src, void* dst, int32_t dst_dtype, int32_t planes, int32_t C, int32_t HW,
                    tc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K3  Linear + fused epilogue.   Replaces every nn.Linear / LayerNorm / ReLU / residual chain on the path:
 * mmcv FFN + norms (cfg detr3d_res101_gridmask.py:65-82), T:362,375,377 (attention_weights, output_proj,
 * position_encoder), T:191 (reg_branches), H:533-536 (radar encoders), H:578 (q/k/v/out projections of
 * nn.MultiheadAttention), H:583-586 (rf_norm2, rf_linear1/2, rf_norm3), H:592-593 (final_cls/final_reg).
 *
 *   Y = A[M,K] * W[N,K]^T                    (fp32 accumulate)
 *   Y += bias[n]                              if bias
 *   Y += row_bias[m % row_bias_period, n]     if row_bias     (e.g. the input-independent query_pos * W^T term)
 *   Y  = row_gate[m] ? Y : 0                  if row_gate     (rows whose attention saw no key: quirk Q6)
 *   Y += residual[m, n] (+ residual2[m, n])   if residual(2)
 *   Y  = LayerNorm_N(Y) * gamma + beta        if ln_gamma     (eps = ln_eps, biased variance; needs N <= 512)
 *   Y  = max(Y, 0)                            if relu
 *   Y += post_add[m, n]                       if post_add     (H:536: pos_feat + ReLU(feat))
 *   out_f32[m, n] = Y ; out_bf16[m, n] = 16-bit copy of Y in out16_dtype    (either may be NULL, not both)
 *
 * a_dtype/w_dtype: both TC_F32 -> exact-fp32 SIMT path (parity mode);  both TC_BF16 -> tcgen05 tensor-core
 * path, one MMA pass (16-byte aligned rows, K >= 64) with SIMT fallback for the tiny odd shapes;  both TC_BF16X2 ->
 * tcgen05 "bf16x3" path: A is [M, 2K], W is [N, 2K] split bf16 (lda / ldw >= 2K, K % 64 == 0), three MMA passes
 * (hi*hi + lo*hi + hi*lo) into the same fp32 accumulator.
 * tail (optional, N <= 32, no LayerNorm): a row-local stage fused behind the epilogue of the last Linear of a regression
 * branch, so that the refinement / anchor / mask-geometry steps cost no launches of their own:
 *   TC_TAIL_REF_UPDATE (T:195-203)  tail_ref_out[m, 0:3] = sigmoid(Y[m, {0,1,4}] + inverse_sigmoid(tail_in[m, 0:3]));
 *                                   if tail_geom_out: the radar-mask geometry of H:543-567 from the NEW reference point
 *                                   (x, y mapped to metres with tail_pc_range) and Y as the box code
 *   TC_TAIL_BOX (H:596-600, :664-665, :722-723)  Y[m, 0:2] += anchor xy, Y[m, 4] += anchor z before Y is stored, anchor =
 *                                   tail_in[m, {tail_xy_col, tail_xy_col + 1, tail_z_col}] (x, y mapped from [0,1] to metres
 *                                   when tail_from_norm; z added as is: quirk Q3);  if tail_geom_out: the NEXT radar layer's
 *                                   geometry (H:615-635 / H:671-693) from the updated Y
 *   tail_geom_out rows are tc_radar_geometry's (cx, cy, fx, fy, rx, ry, radius, thr) with the clamp [tail_r_lo, tail_r_hi];
 *   the arithmetic is the same device function the stand-alone kernels use (bit-identical masks).
 * out16_dtype selects the 16-bit output format: 0 or TC_BF16 -> bf16 [M, N];  TC_BF16X2 -> split bf16 [M, 2N]
 * (hi at column n, lo at column N + n; ld_out_bf16 >= 2N) - the A operand of the next bf16x3 Linear;  TC_F16 -> IEEE
 * half [M, N], saturated to +-65504 - the q/k/v operands of the dense attention core in bf16x3 mode.
 * w_static != 0 promises that W is not written by any earlier launch still in flight on the stream (a weight, not an
 * activation): the tensor-core kernel then fetches the first W tiles BEFORE it waits for its predecessor (programmatic
 * dependent launch), hiding their latency behind the previous kernel's tail.  Leave 0 when W is produced on the stream
 * (the dgrad / wgrad GEMMs of the training variant).
 */
enum { TC_TAIL_NONE = 0, TC_TAIL_REF_UPDATE = 1, TC_TAIL_BOX = 2 };
typedef struct {
  const void* A;  int32_t a_dtype;  int64_t lda;      /* elements */
  const void* W;  int32_t w_dtype;  int64_t ldw;
  int32_t M, N, K;
  const float* bias;
  const float* row_bias;  int32_t row_bias_period;  int64_t ld_row_bias;
  const uint8_t* row_gate;
  const float* residual;  int64_t ld_residual;
  const float* residual2; int64_t ld_residual2;
  const float* ln_gamma;  const float* ln_beta;  float ln_eps;
  int32_t relu;
  const float* post_add;  int64_t ld_post_add;
  float* out_f32;  int64_t ld_out_f32;
  void*  out_bf16; int64_t ld_out_bf16;
  int32_t out16_dtype;
  int32_t tail;                                   /* TC_TAIL_NONE / TC_TAIL_REF_UPDATE / TC_TAIL_BOX */
  const float* tail_in; int64_t ld_tail_in;
  float* tail_ref_out;                            /* [M, 3] contiguous (TC_TAIL_REF_UPDATE) */
  float* tail_geom_out;                           /* [M, 8] contiguous, optional */
  int32_t tail_xy_col, tail_z_col, tail_from_norm;
  float tail_pc_range[6];
  float tail_r_lo, tail_r_hi;
  int32_t w_static;                               /* W is a parameter no earlier launch of the stream writes (see above) */
} tc_linear_args;
TC_API int tc_linear(const tc_linear_args* a, tc_stream_t stream);

/* Fused feed-forward block in the bf16x3 operand format:  Y = LayerNorm(residual + W2 relu(W1 X + b1) + b2).
 * Replaces the mmcv FFN + norm of a decoder layer (BaseTransformerLayer 'ffn', 'norm': ffns.0.layers.0.0 / layers.1 +
 * norms.2, configured at projects/configs/detr3d/detr3d_res101_gridmask.py:65-82: feedforward_channels = 512, operation_order
 * (..., 'ffn', 'norm')) and H:583-586 / H:655-658 / H:713-716 (rf_linear2(relu(rf_linear1(x))) + rf_norm3; modules at H:131-137) - the same result as two tc_linear calls (relu, then residual + LayerNorm), as one launch
 * in which the [M, H] hidden activation never leaves the SM pair that produced it (csrc/ffn_tc.cu).
 * X [M, 2C], W1 [H, 2C], W2 [C, 2H] are split bf16 (TC_BF16X2: hi | lo), residual fp32 [M, C] (the identity branch, normally
 * X in fp32).  Outputs: out_f32 [M, C] and / or out16 split bf16 [M, 2C]; either may be NULL.  Built for C = 256, H = 512 (the
 * TransCAR configs); other sizes return TC_ERR_SHAPE and callers fall back to two tc_linear calls.  w_static: as in tc_linear. */
typedef struct {
  const void* X; int64_t ldx;
  const void* W1; int64_t ldw1; const float* b1;
  const void* W2; int64_t ldw2; const float* b2;
  const float* residual; int64_t ld_residual;
  const float* ln_gamma; const float* ln_beta; float ln_eps;
  float* out_f32; int64_t ld_out_f32;
  void* out16; int64_t ld_out16;
  int32_t M, C, H;
  int32_t w_static;
} tc_ffn_args;
TC_API int tc_ffn(const tc_ffn_args* a, tc_stream_t stream);

/* Fused three-layer head in the bf16x3 operand format:  Y = W3 f2(W2 f1(W1 X + b1) + b2) + b3 with f = ReLU, or
 * ReLU(LayerNorm(.)) when that layer's ln gamma / beta are given, followed by the row-local tail of tc_linear
 * (TC_TAIL_REF_UPDATE / TC_TAIL_BOX on the first 8 output columns; same tail_* fields).  Replaces the three tc_linear launches
 * of reg_branches[l] (Linear-ReLU-Linear-ReLU-Linear, built at H:208-213 / H:224-225; used at T:190-203), final_reg* (built
 * at H:84-90 / H:102-108 / H:120-126; used with the anchor add at H:593-600, H:662-665, H:720-723) and final_cls*
 * (Linear-LN-ReLU-Linear-LN-ReLU-Linear, H:74-83 / H:92-101 / H:110-119; used at H:592, H:661, H:719): the two hidden activations stay in tensor /
 * shared memory of the CTA that owns the 128 rows (csrc/mlp_tc.cu).  X [M, 2C], W1 / W2 [C, 2C], W3 [N3, 2C] are split bf16;
 * out_f32 [M, N3] fp32.  Built for C = 256, N3 <= 32; other sizes return TC_ERR_SHAPE (callers use three tc_linear calls). */
typedef struct {
  const void* X; int64_t ldx;
  const void* W1; int64_t ldw1; const float* b1; const float* ln1_gamma; const float* ln1_beta;
  const void* W2; int64_t ldw2; const float* b2; const float* ln2_gamma; const float* ln2_beta;
  const void* W3; int64_t ldw3; const float* b3;
  float ln_eps;
  float* out_f32; int64_t ld_out_f32;
  int32_t M, C, N3;
  int32_t w_static;
  int32_t tail;
  const float* tail_in; int64_t ld_tail_in;
  float* tail_ref_out; float* tail_geom_out;
  int32_t tail_xy_col, tail_z_col, tail_from_norm;
  float tail_pc_range[6];
  float tail_r_lo, tail_r_hi;
} tc_mlp_args;
TC_API int tc_mlp(const tc_mlp_args* a, tc_stream_t stream);

/* Fused 3 -> C position encoder head: Y = ReLU(LayerNorm(Linear_{3->C}(f(x)))), f = inverse_sigmoid (eps 1e-5,
 * T:17-32) when logit_input != 0 else identity.  Replaces T:377 (position_encoder[0:3]) and H:533
 * (radar_position_encoder[0:3]).  x is [M, ldx] fp32 (first 3 columns used); C <= 1024, C % 32 == 0. */
typedef struct {
  const float* x; int64_t ldx; int32_t M; int32_t C; int32_t logit_input;
  const float* weight;   /* [C,3] */
  const float* bias;     /* [C]   */
  const float* ln_gamma; const float* ln_beta; float ln_eps;
  float* out_f32; void* out_bf16;   /* [M,C]; either may be NULL */
  int32_t out16_dtype;              /* 0 / TC_BF16: out_bf16 is bf16 [M,C];  TC_BF16X2: split bf16 [M,2C] */
} tc_point_embed_args;
TC_API int tc_point_embed(const tc_point_embed_args* a, tc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K4  multi-head attention core softmax(scale * Q K^T + mask) V on already-projected operands.
 * Replaces the baddbmm -> softmax -> bmm core of nn.MultiheadAttention (slow path) for the
 * decoder self-attention (mask-free) and for H:578 / H:649 / H:707 with the radar distance mask of
 * H:549-571 built IN-KERNEL from per-query geometry (no [Q,R] mask tensor in HBM, no torch.where sync).
 *
 *   q [B, Lq, heads*D]  k, v [B, Lk, heads*D]   (dtype qkv_dtype: fp32, bf16 or fp16; ld* = row stride in elements)
 *   out [B, Lq, heads*D] out_dtype (fp32, bf16, or TC_BF16X2: split bf16 [B, Lq, 2*heads*D], hi | lo, ldo >= 2*heads*D)
 *   geom (optional) [B, Lq, 8] fp32 = (cx, cy, fx, fy, rx, ry, radius, thr): centre / front / rear circle
 *         centres in metres, the clamped radius and thr = the smallest fp32 whose square root is >= radius
 *         (sqrt(x) < radius  <=>  x < thr) - produced by tc_radar_geometry
 *   key_xy (with geom) [B, Lk, 2] fp32 radar x,y in metres (padding slots hold 500)
 *   row_any (optional) [B, Lq] uint8: 1 if the row has at least one allowed key.  Rows without any
 *         allowed key produce out = 0 (they skip attention in the reference: quirk Q6).
 *   algo  which kernel runs (tc_attention_algo).  The TransCAR mask admits ~0.1 % of the (query, radar point)
 *         pairs, so TC_ATTN_AUTO sends masked calls to the sparse kernel (all 8 heads share one key scan and
 *         only allowed pairs are computed) and mask-free calls (decoder self-attention) to the tcgen05 kernel;
 *         an explicit value that the shape/dtype does not support is TC_ERR_SHAPE / TC_ERR_DTYPE.
 */
typedef enum {
  TC_ATTN_AUTO = 0,     /* masked + 8x32 heads -> TC_ATTN_SPARSE; bf16 / fp16 dense -> TC_ATTN_TENSOR; else TC_ATTN_SIMT */
  TC_ATTN_TENSOR = 1,   /* tcgen05/TMA dense-tile kernel (bf16 or fp16 operands), mask evaluated per tile in-kernel */
  TC_ATTN_SIMT = 2,     /* CUDA-core dense-tile kernel (fp32 parity mode) */
  TC_ATTN_SPARSE = 3    /* radar path: per-query key scan, only allowed (query, key) pairs are computed (needs geom) */
} tc_attention_algo;

typedef struct {
  const void* q; const void* k; const void* v;
  int64_t ldq, ldk, ldv;
  int64_t q_batch_stride, k_batch_stride, v_batch_stride;
  int32_t qkv_dtype;
  int32_t B, Lq, Lk, heads, D;
  float scale;
  const float* geom;
  const float* key_xy;
  void* out; int64_t ldo; int32_t out_dtype;
  uint8_t* row_any;
  int32_t algo;                /* tc_attention_algo; TC_ATTN_AUTO (0) picks the fastest exact path */
  /* training variant: dropout on the attention probabilities (nn.MultiheadAttention(dropout=0.1), H:128 / mmcv attn_drop),
   * mask = tc_dropout's for the logical [B*heads*Lq, Lk] tensor, row = (b * heads + h) * Lq + q.  SIMT and sparse paths. */
  float dropout_p; uint64_t dropout_seed; uint64_t dropout_stream;
  /* generic masks of nn.MultiheadAttention (the mmcv wrapper's attn_mask / key_padding_mask arguments; never used by the
   * TransCAR configs, served by the SIMT path): attn_blocked [Lq, Lk] uint8 shared by all samples and heads,
   * key_blocked [B, Lk] uint8; 1 = the key is not attended.  Either may be NULL. */
  const uint8_t* attn_blocked; const uint8_t* key_blocked;
} tc_attention_args;
TC_API int tc_attention_fwd(const tc_attention_args* a, tc_stream_t stream);

/* Per-query circle geometry for the radar mask.  Replaces H:543-567 / H:615-635 / H:671-693.
 *   centre [M, ld_centre] fp32: columns 0,1 = x,y.  centre_is_normalised != 0 -> x,y are in [0,1] and
 *           are mapped to metres with pc_range first (layer 1: inter_references[-1]); else already metres.
 *   code   [M, ld_code] fp32: previous stage regression; columns 3 (log length), 6, 7 (heading terms) used.
 *   geom   [M, 8] fp32 out.   radius = clamp(exp(code[3]) / 2, r_lo, r_hi)  (quirk Q7). */
typedef struct {
  const float* centre; int64_t ld_centre; int32_t centre_is_normalised;
  const float* code;   int64_t ld_code;
  int32_t M;
  float pc_range[6];
  float r_lo, r_hi;
  float* geom;
} tc_radar_geometry_args;
TC_API int tc_radar_geometry(const tc_radar_geometry_args* a, tc_stream_t stream);

/* Materialise the [B, Lq, Lk] uint8 "blocked" mask (1 = key not attended) with exactly the arithmetic the
 * attention kernel uses in-kernel: torch.cdist's mm-based Euclidean distance, '<' against the
 * radius, OR of the three circles, NOT.  Test / debugging aid; the hot path never materialises it. */
TC_API int tc_radar_mask(const float* geom, const float* key_xy, int32_t B, int32_t Lq, int32_t Lk,
                  uint8_t* blocked, uint8_t* row_any, tc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Small fused pointwise stages.
 */
/* Iterative reference refinement T:195-203:  new.xy = sigmoid(code.xy + logit(ref.xy)),
 * new.z = sigmoid(code[4] + logit(ref.z)).   code [M, ld_code], ref/new_ref [M,3]. */
TC_API int tc_ref_update(const float* code, int64_t ld_code, const float* ref, float* new_ref, int32_t M,
                  tc_stream_t stream);

/* Box-code assembly H:596-600 / H:664-665 / H:722-723:  code[0:2] += anchor_xy, code[4] += anchor_z, in place.
 * anchor [M, ld_anchor]; xy_col/z_col select the columns; xy_from_normalised != 0 maps anchor x,y in [0,1]
 * to metres with pc_range first while z is added as is (quirk Q3). */
TC_API int tc_box_anchor_add(float* code, int64_t ld_code, const float* anchor, int64_t ld_anchor, int32_t xy_col,
                      int32_t z_col, int32_t xy_from_normalised, const float* pc_range6, int32_t M,
                      tc_stream_t stream);

/* fp32 -> bf16 cast of a [rows, cols] matrix (weights are cast once, activations by fused epilogues). */
TC_API int tc_cast_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int32_t rows, int32_t cols,
                 tc_stream_t stream);

/* fp32 [rows, cols] -> split bf16 [rows, 2 * cols] (TC_BF16X2: hi | lo); ld_dst >= 2 * cols. */
TC_API int tc_cast_split(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int32_t rows, int32_t cols,
                  tc_stream_t stream);

/* N1  NMS-free decode on device.  Replaces C:39-90 + U:26-52: sigmoid, top-`max_num` over Q*classes,
 * denormalise (exp sizes, atan2 heading), centre-range test.  Fixed-size outputs, no host sync:
 *   boxes [B, max_num, 9], scores [B, max_num], labels [B, max_num] int32, keep [B, max_num] uint8, and / or
 *   records [B, max_num, 12] fp32 = (9 box values, score, label, keep) - the fixed-size per-sample record that the
 *   multi-GPU result gather ships (replaces mmdet's collect_results pickling, tools/test.py:218-223).
 *   Either output group may be NULL (boxes / scores / labels / keep go together).
 * workspace: tc_decode_workspace_bytes(B, Q, classes) bytes. */
typedef struct {
  const float* cls; const float* code;       /* [B,Q,classes], [B,Q,10] */
  int32_t B, Q, classes, max_num;
  float post_center_range[6];
  float* boxes; float* scores; int32_t* labels; uint8_t* keep;
  void* workspace;
  float* records;
} tc_decode_args;
TC_API int64_t tc_decode_workspace_bytes(int32_t B, int32_t Q, int32_t classes);
TC_API int tc_decode(const tc_decode_args* a, tc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Training variant: backward building blocks of the radar fusion head - the part of TransCAR that trains
 * (tools/train.py:238-252 freezes backbone, neck, DETR3D transformer, cls/reg branches and query embedding, so
 * gradients flow through H:531-536 and H:573-729 only).  The backward GEMMs are tc_linear calls on transposed
 * operands (dX = dY W: A = dY, W = W^T;  dW = dY^T X: A = dY^T, W = X^T) - on the tensor cores in bf16x3 mode, the
 * transposes written directly as split-bf16 operands by tc_transpose; everything else is below.  Gradients that
 * are reductions (bias / LayerNorm parameters, dK / dV rows shared by many queries) are ACCUMULATED with fp32
 * atomics: the caller zeroes its gradient bucket once per step.
 */
/* dst[c, r] = src[r, c]; dtypes fp32 or bf16 on either side (ld in elements).  dst_dtype = TC_BF16X2 (fp32 source only)
 * writes the split-bf16 transpose [cols, 2 * rows_pad] with rows_pad = ld_dst / 2 >= rows; columns rows..rows_pad-1 of both
 * halves are zero-filled, so rows_pad can be the 64-aligned reduction length of a bf16x3 wgrad GEMM. */
TC_API int tc_transpose(const void* src, int32_t src_dtype, int64_t ld_src, void* dst, int32_t dst_dtype, int64_t ld_dst,
                 int32_t rows, int32_t cols, tc_stream_t stream);
/* out[n] += sum_m x[m, n]   (bias gradient). */
TC_API int tc_colsum(const void* x, int32_t dtype, int64_t ldx, int32_t M, int32_t N, float* out, tc_stream_t stream);

/* y = LayerNorm_N(x) * gamma + beta (biased variance, eps), optional ReLU; saves mean / rstd for the backward.
 * Replaces nn.LayerNorm where the training forward keeps it un-fused (rf_norm2/3*, final_cls*.1/.4,
 * radar_position_encoder.1/.4). */
typedef struct {
  const float* x; int64_t ldx; int32_t M, N;
  const float* gamma; const float* beta; float eps; int32_t relu;
  float* y_f32; void* y_bf16; int64_t ldy;       /* either output may be NULL */
  float* mean; float* rstd;                      /* [M], may be NULL */
} tc_layernorm_args;
TC_API int tc_layernorm_fwd(const tc_layernorm_args* a, tc_stream_t stream);

/* dx = LayerNorm backward of dy (+ add, the gradient arriving through a residual connection);
 * dgamma[n] += sum_m dy * xhat, dbeta[n] += sum_m dy. */
typedef struct {
  const float* dy; int64_t ld_dy;
  const float* x; int64_t ldx;
  const float* mean; const float* rstd; const float* gamma;
  int32_t M, N;
  const float* add; int64_t ld_add;              /* optional */
  float* dx; int64_t ld_dx;
  float* dgamma; float* dbeta;                   /* optional, accumulated */
} tc_layernorm_bwd_args;
TC_API int tc_layernorm_bwd(const tc_layernorm_bwd_args* a, tc_stream_t stream);

/* dz = dy * (y > 0 if y) * (gate[m] != 0 if gate): ReLU backward and / or the attention row gate (quirk Q6). */
TC_API int tc_mask_grad(const float* dy, int64_t ld_dy, const float* y, int64_t ld_y, const uint8_t* gate, float* dz,
                 int64_t ld_dz, int32_t M, int32_t N, tc_stream_t stream);

/* Backward of the masked radar attention core (H:578) for fp32 operands: dq is written, dk / dv are accumulated
 * (one radar point is attended by several queries).  Same key scan and bit-exact mask as the forward sparse kernel;
 * the softmax is recomputed, nothing but q, k, v needs to be kept from the forward. */
typedef struct {
  const float* q; const float* k; const float* v; const float* dout;
  int64_t ldq, ldk, ldv, ld_dout;
  int64_t q_batch_stride, k_batch_stride, v_batch_stride;
  int32_t B, Lq, Lk, heads, D;
  float scale;
  const float* geom; const float* key_xy;
  float* dq; int64_t ld_dq;                      /* [B*Lq, heads*D] */
  float* dk; float* dv; int64_t ld_dk, ld_dv, dk_batch_stride, dv_batch_stride;
  float dropout_p; uint64_t dropout_seed; uint64_t dropout_stream;     /* as in the forward call */
} tc_attention_bwd_args;
TC_API int tc_attention_sparse_bwd(const tc_attention_bwd_args* a, tc_stream_t stream);

/* Backward of K1 (the "grid_sample scatter" of the training variant).  Replaces autograd through T:367-373 + T:381-422:
 * F.grid_sample backward (bilinear, zeros padding, align_corners=False) for the four levels, the sigmoid-weighted masked
 * sum and the projection chain, in one kernel over the same (query, valid camera, level, corner) set as the forward.
 *   feat / ref / lidar2img / attn_logits / pc_range / img_w / img_h: exactly the forward's inputs (tc_sample_args)
 *   dout      [B, Q, C] fp32         upstream gradient of the forward's `out`
 *   d_feat[l] [B, N, H_l, W_l, C] fp32 channels-last, optional (all four or none): ACCUMULATED with fp32 atomics
 *   d_logits  [B, Q, N*L] fp32, optional: written (0 for cameras that fail the validity test)
 *   d_ref     [B, Q, 3] fp32, optional: written; gradient w.r.t. the normalised reference points (layer 0 of the decoder is
 *             the only layer whose reference points are not detached, T:203)
 */
typedef struct {
  const void* feat[TC_MAX_LEVELS];
  int32_t H[TC_MAX_LEVELS];
  int32_t W[TC_MAX_LEVELS];
  int32_t num_levels;
  int32_t B, N, Q, C;
  int32_t feat_dtype;
  const float* ref;
  const float* lidar2img;
  const float* attn_logits;
  float pc_range[6];
  float img_w, img_h;
  const float* dout;
  float* d_feat[TC_MAX_LEVELS];
  float* d_logits;
  float* d_ref;
} tc_sample_bwd_args;
TC_API int tc_sample_bwd(const tc_sample_bwd_args* a, tc_stream_t stream);

/* Backward of the dense (mask-free) attention core of the decoder self-attention, fp32 operands: softmax statistics are
 * recomputed from q, k (nothing but q, k, v, o is kept from the forward).  Replaces autograd through the baddbmm ->
 * softmax -> bmm core of nn.MultiheadAttention (mmcv wrapper of cfg detr3d_res101_gridmask.py:68-72).
 *   q, o, dout [B, Lq, heads*D]; k, v [B, Lk, heads*D] (row strides ld*, batch strides *_batch_stride, elements)
 *   dq [B, Lq, heads*D], dk / dv [B, Lk, heads*D]: written (contiguous, row stride heads*D)
 *   workspace: 2 * B * heads * Lq floats (log-sum-exp and <dout, o> per query row and head)
 */
typedef struct {
  const float* q; const float* k; const float* v; const float* o; const float* dout;
  int64_t ldq, ldk, ldv, ldo, ld_dout;
  int64_t q_batch_stride, k_batch_stride, v_batch_stride, o_batch_stride, dout_batch_stride;
  int32_t B, Lq, Lk, heads, D;
  float scale;
  float* dq; float* dk; float* dv;
  float* workspace;
  float dropout_p; uint64_t dropout_seed; uint64_t dropout_stream;     /* as in the forward call */
} tc_attention_dense_bwd_args;
TC_API int tc_attention_dense_bwd(const tc_attention_dense_bwd_args* a, tc_stream_t stream);

/* Pointwise pieces of the decoder's training variant (T:17-32 inverse_sigmoid, T:122-123 / T:195-203 sigmoid), n elements:
 *   TC_PW_LOGIT_BWD    out = grad * d inverse_sigmoid(x) / dx   (torch.clamp gradients: zero outside [0,1] / below eps)
 *   TC_PW_SIGMOID_BWD  out = grad * y (1 - y)                   (x = the sigmoid OUTPUT y)
 *   TC_PW_LOGIT        out = inverse_sigmoid(x)                 (grad ignored, may be NULL)
 *   TC_PW_SIGMOID      out = sigmoid(x)                         (grad ignored, may be NULL) */
enum { TC_PW_LOGIT_BWD = 0, TC_PW_SIGMOID_BWD = 1, TC_PW_LOGIT = 2, TC_PW_SIGMOID = 3 };
TC_API int tc_pointwise(const float* grad, const float* x, float* out, int32_t n, int32_t mode, tc_stream_t stream);

/* Dropout with a regenerable mask (training variant; rf_dropout* H:133-145, mmcv proj / ffn dropout):
 *   out[m, n] = (residual ? residual[m, n] : 0) + keep(m, n) * x[m, n] / (1 - p),   keep ~ Bernoulli(1 - p)
 * keep is a pure function of (seed, stream, m, n) (Philox4x32-10), so the backward pass applies the SAME call to the
 * gradient instead of storing a mask.  x, residual, out: contiguous fp32 [M, N]; out may alias x or residual. */
TC_API int tc_dropout(const float* x, const float* residual, float* out, int32_t M, int32_t N, float p, uint64_t seed,
               uint64_t stream, tc_stream_t s);

/* out[m, :] = a[m, :] + b[m % period, :]   ([M, N] fp32, contiguous; out may alias a).  The query_pos broadcast of the
 * decoder's training forward (q = k = x + query_pos) and gradient sums. */
TC_API int tc_add_rows(const float* a, const float* b, float* out, int32_t M, int32_t N, int32_t period, tc_stream_t stream);
/* out[r, :] += sum_b x[b * period + r, :]   (x [B*period, N] -> out [period, N]): gradients of batch-broadcast parameters
 * (query embedding, T:119-121). */
TC_API int tc_period_sum(const float* x, float* out, int32_t batches, int32_t period, int32_t N, tc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * N4  loss side (training variant).  Problems are indexed p = layer * B + sample over the `layers` output layers.
 *   gt_boxes   [sumG, 9] fp32  (cx, cy, cz, w, l, h, rot, vx, vy), gravity centre (H:962-964), all samples concatenated
 *   gt_labels  [sumG] int32;   gt_offsets [B + 1] int32: sample b owns rows gt_offsets[b] .. gt_offsets[b+1]-1
 *
 * tc_match_cost: cost [layers*B, Q, Gmax] fp32 = cls_weight * FocalLossCost(cls, label) + reg_weight * sum_j |bbox_j -
 * normalize_bbox(gt)_j| (hungarian_assigner_3d.py:106-115, match_cost.py:15-26); columns >= G_b hold +inf. */
typedef struct {
  const float* cls; const float* bbox;            /* [layers*B, Q, classes], [layers*B, Q, 10] */
  const float* gt_boxes; const int32_t* gt_labels; const int32_t* gt_offsets;
  int32_t layers, B, Q, classes, Gmax;
  float cls_weight, reg_weight, alpha, gamma, eps;
  float* cost;
} tc_match_cost_args;
TC_API int tc_match_cost(const tc_match_cost_args* a, tc_stream_t stream);

/* tc_detr_loss: sigmoid focal loss (mmdet FocalLoss, use_sigmoid, label weight 1) + code-weighted L1 loss on the matched
 * rows (H:849-917) of every layer, with their gradients.
 *   assigned  [layers*B, Q] int32: -1 = background, else the row of gt_boxes / gt_labels matched to the query
 *   cls_avg / pos_avg [layers] fp32: the (already rank-averaged, >= 1) normalisers of H:885-897
 *   loss_cls / loss_bbox [layers] fp32: ACCUMULATED (zero them first);  d_cls [layers*B, Q, classes], d_bbox [.., 10]: written
 *   (may be NULL).  Rows whose normalised target is not finite carry no box loss (H:899). */
typedef struct {
  const float* cls; const float* bbox;
  const int32_t* assigned;
  const float* gt_boxes; const int32_t* gt_labels;
  const float* code_weights;                      /* [10] */
  const float* cls_avg; const float* pos_avg;
  int32_t layers, B, Q, classes;
  float alpha, gamma, loss_cls_weight, loss_bbox_weight;
  float* loss_cls; float* loss_bbox;
  float* d_cls; float* d_bbox;
} tc_detr_loss_args;
TC_API int tc_detr_loss(const tc_detr_loss_args* a, tc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif  /* TRANSCAR_B200_H_ */
