// Backward of K1 (fused camera sampling): the grid_sample scatter of BASELINE.json configs[4].
//
// Forward (sample.cu):  out[b,q,:] = sum_{cam valid} sum_{level} sigmoid(logit[b,q,cam*L+level]) * sum_{corner} w_corner * texel
// Given dout [B,Q,C] this kernel produces, in one pass over the same (query, camera, level, corner) set:
//   d_feat[level][texel,:] += sigmoid * w_corner * dout          fp32 atomics (red.global.add.v4.f32) into channels-last maps
//   d_logits[b,q,cam*L+level] = sigmoid (1 - sigmoid) * <dout, bilinear sample of the level>
//   d_ref[b,q,0:3]            = chain rule through the bilinear weights (ATen grid_sampler_2d backward: out-of-range corners
//                               contribute nothing), the un-normalisation, the perspective divide and lidar2img
//                               (detr3d_transformer.py:389-411).  The validity mask and the eps clamp are not differentiable;
//                               a valid camera always has depth > eps, so the clamp is inactive wherever gradient flows.
// One warp per (sample, query); lane l owns channels [8l, 8l+8) of a 256-channel chunk, exactly like the forward consume step.
// This is a training-variant kernel (the reference recipe freezes the decoder): written for correctness and coalesced
// atomics, not tuned to the roofline.
#include "tc_common.cuh"

namespace tc {
namespace {

struct SampleBwdParams {
  const void* feat[TC_MAX_LEVELS];
  float* d_feat[TC_MAX_LEVELS];
  int H[TC_MAX_LEVELS];
  int W[TC_MAX_LEVELS];
  int B, N, Q, C;
  const float* ref;
  const float* lidar2img;
  const float* logits;
  const float* dout;
  float pc[6];
  float img_w, img_h;
  float* d_logits;
  float* d_ref;
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool kBf16>
__global__ void __launch_bounds__(256) sample_bwd_kernel(const SampleBwdParams p) {
  const int lane = threadIdx.x & 31;
  const int task = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (task >= p.B * p.Q) return;
  const int b = task / p.Q;
  const float rx = p.ref[(size_t)task * 3 + 0], ry = p.ref[(size_t)task * 3 + 1], rz = p.ref[(size_t)task * 3 + 2];
  const float sx = p.pc[3] - p.pc[0], sy = p.pc[4] - p.pc[1], sz = p.pc[5] - p.pc[2];
  const float px = __fadd_rn(__fmul_rn(rx, sx), p.pc[0]);
  const float py = __fadd_rn(__fmul_rn(ry, sy), p.pc[1]);
  const float pz = __fadd_rn(__fmul_rn(rz, sz), p.pc[2]);
  const float inv_w = 1.0f / p.img_w, inv_h = 1.0f / p.img_h;
  float dref[3] = {0.f, 0.f, 0.f};

  for (int cam = 0; cam < p.N; ++cam) {
    const float* M = p.lidar2img + ((size_t)b * p.N + cam) * 16;
    // forward projection, identical arithmetic to sample.cu (the validity decision must not differ)
    const float cx = __fadd_rn(__fmaf_rn(M[1], py, __fmul_rn(M[0], px)), __fmaf_rn(M[3], 1.0f, __fmul_rn(M[2], pz)));
    const float cy = __fadd_rn(__fmaf_rn(M[5], py, __fmul_rn(M[4], px)), __fmaf_rn(M[7], 1.0f, __fmul_rn(M[6], pz)));
    const float cz = __fadd_rn(__fmaf_rn(M[9], py, __fmul_rn(M[8], px)), __fmaf_rn(M[11], 1.0f, __fmul_rn(M[10], pz)));
    const float eps = 1e-5f;
    const float zc = fmaxf(cz, eps);
    const float u = __fmul_rn(__fdiv_rn(cx, zc), inv_w), v = __fmul_rn(__fdiv_rn(cy, zc), inv_h);
    const float gx = __fmul_rn(__fadd_rn(u, -0.5f), 2.0f), gy = __fmul_rn(__fadd_rn(v, -0.5f), 2.0f);
    const bool valid = (cz > eps) && (gx > -1.0f) && (gx < 1.0f) && (gy > -1.0f) && (gy < 1.0f);
    if (!valid) {                                                    // masked out: no gradient anywhere
      if (p.d_logits && lane < 4) p.d_logits[(size_t)task * (p.N * 4) + cam * 4 + lane] = 0.f;
      continue;
    }
    float d_gx = 0.f, d_gy = 0.f;                                    // d loss / d grid coordinate, summed over levels
    for (int level = 0; level < 4; ++level) {
      const int H = p.H[level], W = p.W[level];
      const float logit = p.logits[(size_t)task * (p.N * 4) + cam * 4 + level];
      const float sig = sigmoid_f32(logit);
      const float ix = (__fadd_rn(gx, 1.0f) * (float)W - 1.0f) * 0.5f, iy = (__fadd_rn(gy, 1.0f) * (float)H - 1.0f) * 0.5f;
      const float fx0 = floorf(ix), fy0 = floorf(iy);
      float dot_w = 0.f, dot_dx = 0.f, dot_dy = 0.f;                 // sum_corner {w, dw/dix, dw/diy} * <dout, texel>
      for (int corner = 0; corner < 4; ++corner) {
        const int xi = (int)fx0 + (corner & 1), yi = (int)fy0 + (corner >> 1);
        if (xi < 0 || xi >= W || yi < 0 || yi >= H) continue;        // zeros padding (warp-uniform)
        const float wx = (corner & 1) ? ix - fx0 : (fx0 + 1.0f) - ix;
        const float wy = (corner >> 1) ? iy - fy0 : (fy0 + 1.0f) - iy;
        const float dwx = (corner & 1) ? 1.0f : -1.0f, dwy = (corner >> 1) ? 1.0f : -1.0f;
        const size_t texel = (((size_t)b * p.N + cam) * H + yi) * W + xi;
        float acc = 0.f;
        for (int c0 = lane * 8; c0 < p.C; c0 += 256) {
          float t[8], g[8];
          const float4* gp = reinterpret_cast<const float4*>(p.dout + (size_t)task * p.C + c0);
          const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1);
          g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
          if (kBf16) {
            const uint4 q4 = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.feat[level]) + texel * p.C + c0));
            t[0] = bf16_lo(q4.x); t[1] = bf16_hi(q4.x); t[2] = bf16_lo(q4.y); t[3] = bf16_hi(q4.y);
            t[4] = bf16_lo(q4.z); t[5] = bf16_hi(q4.z); t[6] = bf16_lo(q4.w); t[7] = bf16_hi(q4.w);
          } else {
            const float4* tp = reinterpret_cast<const float4*>(static_cast<const float*>(p.feat[level]) + texel * p.C + c0);
            const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
            t[0] = t0.x; t[1] = t0.y; t[2] = t0.z; t[3] = t0.w; t[4] = t1.x; t[5] = t1.y; t[6] = t1.z; t[7] = t1.w;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) acc = fmaf(g[i], t[i], acc);
          if (p.d_feat[level]) {
            const float s = sig * wx * wy;
            float* dp = p.d_feat[level] + texel * p.C + c0;
            red_add_v4(dp, s * g[0], s * g[1], s * g[2], s * g[3]);
            red_add_v4(dp + 4, s * g[4], s * g[5], s * g[6], s * g[7]);
          }
        }
        const float dot = warp_sum(acc);
        dot_w = fmaf(wx * wy, dot, dot_w);
        dot_dx = fmaf(dwx * wy, dot, dot_dx);
        dot_dy = fmaf(wx * dwy, dot, dot_dy);
      }
      if (p.d_logits && lane == 0) p.d_logits[(size_t)task * (p.N * 4) + cam * 4 + level] = sig * (1.0f - sig) * dot_w;
      d_gx = fmaf(sig * dot_dx, 0.5f * (float)W, d_gx);              // d ix / d gx = W / 2
      d_gy = fmaf(sig * dot_dy, 0.5f * (float)H, d_gy);
    }
    // grid -> image plane -> camera frame -> lidar frame -> normalised reference point
    const float d_u = d_gx * 2.0f * inv_w, d_v = d_gy * 2.0f * inv_h;         // u = cx / zc / img_w,  gx = (u - 0.5) * 2
    const float d_cx = d_u / zc, d_cy = d_v / zc;
    const float d_cz = -(d_u * cx + d_v * cy) / (zc * zc);
    dref[0] += (M[0] * d_cx + M[4] * d_cy + M[8] * d_cz) * sx;
    dref[1] += (M[1] * d_cx + M[5] * d_cy + M[9] * d_cz) * sy;
    dref[2] += (M[2] * d_cx + M[6] * d_cy + M[10] * d_cz) * sz;
  }
  if (p.d_ref && lane < 3) p.d_ref[(size_t)task * 3 + lane] = dref[lane];
}

}  // namespace
}  // namespace tc

extern "C" int tc_sample_bwd(const tc_sample_bwd_args* a, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(a != nullptr, TC_ERR_NULL, "tc_sample_bwd: args is NULL");
  if (a->B == 0 || a->Q == 0) return TC_OK;
  TC_REQUIRE(a->ref && a->lidar2img && a->attn_logits && a->dout, TC_ERR_NULL, "tc_sample_bwd: NULL tensor pointer");
  TC_REQUIRE(a->num_levels == 4, TC_ERR_SHAPE, "tc_sample_bwd: num_levels must be 4 (got %d)", a->num_levels);
  TC_REQUIRE(a->N >= 1 && a->N <= TC_MAX_CAMS, TC_ERR_SHAPE, "tc_sample_bwd: num cams %d unsupported", a->N);
  TC_REQUIRE(a->C > 0 && a->C % 256 == 0, TC_ERR_SHAPE, "tc_sample_bwd: C must be a multiple of 256 (got %d)", a->C);
  TC_REQUIRE(a->B > 0 && a->Q > 0, TC_ERR_SHAPE, "tc_sample_bwd: bad B/Q");
  TC_REQUIRE(a->feat_dtype == TC_F32 || a->feat_dtype == TC_BF16, TC_ERR_DTYPE, "tc_sample_bwd: bad feat dtype");
  TC_REQUIRE(aligned16(a->dout), TC_ERR_ALIGN, "tc_sample_bwd: dout must be 16-byte aligned");
  TC_REQUIRE(a->d_logits || a->d_ref || a->d_feat[0], TC_ERR_NULL, "tc_sample_bwd: no output requested");
  SampleBwdParams p;
  const bool want_feat = a->d_feat[0] != nullptr;
  for (int l = 0; l < 4; ++l) {
    TC_REQUIRE(a->feat[l] != nullptr && aligned16(a->feat[l]), TC_ERR_ALIGN, "tc_sample_bwd: feat[%d] NULL or not 16-byte aligned", l);
    TC_REQUIRE((a->d_feat[l] != nullptr) == want_feat, TC_ERR_NULL, "tc_sample_bwd: d_feat levels must be all given or all NULL");
    TC_REQUIRE(!a->d_feat[l] || aligned16(a->d_feat[l]), TC_ERR_ALIGN, "tc_sample_bwd: d_feat[%d] must be 16-byte aligned", l);
    TC_REQUIRE(a->H[l] > 0 && a->W[l] > 0, TC_ERR_SHAPE, "tc_sample_bwd: level %d has empty extent", l);
    p.feat[l] = a->feat[l]; p.d_feat[l] = a->d_feat[l]; p.H[l] = a->H[l]; p.W[l] = a->W[l];
  }
  p.B = a->B; p.N = a->N; p.Q = a->Q; p.C = a->C;
  p.ref = a->ref; p.lidar2img = a->lidar2img; p.logits = a->attn_logits; p.dout = a->dout;
  for (int i = 0; i < 6; ++i) p.pc[i] = a->pc_range[i];
  p.img_w = a->img_w; p.img_h = a->img_h;
  p.d_logits = a->d_logits; p.d_ref = a->d_ref;
  const long long total = (long long)a->B * a->Q;
  const unsigned grid = (unsigned)((total + 7) / 8);
  if (a->feat_dtype == TC_BF16) sample_bwd_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(p);
  else sample_bwd_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(p);
  count_launch();
  return check_launch("tc_sample_bwd");
}
