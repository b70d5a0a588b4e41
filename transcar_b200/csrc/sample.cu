// K1: fused DETR3D camera sampling (projection + validity mask + 4-level bilinear gather + masked
// sigmoid-weighted reduction).  See include/transcar_b200.h (tc_sample_fwd) for the contract and the
// reference lines it replaces (detr3d_transformer.py:367-373, 381-422).
//
// Mapping: one warp per (sample, query).  Lane c < N projects the query through camera c's lidar2img
// (matrices staged in shared memory, one sample per block row), a ballot yields the valid-camera set,
// and the warp then walks the valid cameras only (~18 % of (query,cam) pairs are valid, 87 %
// of queries see exactly one camera).  For a valid camera every lane owns 8 channels (bf16: one 128-bit
// load per texel; fp32: two) of the channels-last feature vector, so a texel is one fully coalesced
// 512-byte (bf16) or 2x512-byte (fp32) warp read; all 16 texel reads of a camera (4 levels x 4 corners)
// are issued before the first is consumed.  HBM-bound: algorithmic bytes per valid pair =
// 16 * C * sizeof(feat).
#include "tc_common.cuh"

namespace tc {
namespace {

constexpr int kWarpsPerBlock = 4;

struct SampleParams {
  const void* feat[TC_MAX_LEVELS];
  int H[TC_MAX_LEVELS];
  int W[TC_MAX_LEVELS];
  int B, N, Q, C;
  const float* ref;
  const float* lidar2img;
  const float* logits;
  float pc[6];
  float inv_w, inv_h;      // 1.0f / img_w, 1.0f / img_h
  void* out;
  uint8_t* mask;
};

// ATen grid_sampler_unnormalize, align_corners=False: ((g + 1) * size - 1) / 2.
__device__ __forceinline__ float unnormalize(float g, int size) {
  return (__fadd_rn(g, 1.0f) * static_cast<float>(size) - 1.0f) * 0.5f;
}

template <bool kBf16>
struct Texel;

template <>
struct Texel<true> {                  // 8 bf16 channels per lane: channels [8*lane, 8*lane+8)
  uint4 v;
  __device__ __forceinline__ void load(const void* base, size_t texel, int C, int lane, int chunk) {
    const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(base) + texel * C + chunk * 256 + lane * 8;
    v = ldg_nc_u4(p);
  }
  // compiler-level fence: everything issued before stays before, consumers come after
  __device__ __forceinline__ void pin() { asm volatile("" : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w)); }
  __device__ __forceinline__ void fma_into(float (&acc)[8], float w) const {
    acc[0] = fmaf(bf16_lo(v.x), w, acc[0]); acc[1] = fmaf(bf16_hi(v.x), w, acc[1]);
    acc[2] = fmaf(bf16_lo(v.y), w, acc[2]); acc[3] = fmaf(bf16_hi(v.y), w, acc[3]);
    acc[4] = fmaf(bf16_lo(v.z), w, acc[4]); acc[5] = fmaf(bf16_hi(v.z), w, acc[5]);
    acc[6] = fmaf(bf16_lo(v.w), w, acc[6]); acc[7] = fmaf(bf16_hi(v.w), w, acc[7]);
  }
};

template <>
struct Texel<false> {                 // 8 fp32 channels per lane: [4*lane, +4) and [128 + 4*lane, +4)
  uint4 a, b;
  __device__ __forceinline__ void load(const void* base, size_t texel, int C, int lane, int chunk) {
    const float* p = static_cast<const float*>(base) + texel * C + chunk * 256 + lane * 4;
    a = ldg_nc_u4(p);
    b = ldg_nc_u4(p + 128);
  }
  __device__ __forceinline__ void pin() {
    asm volatile("" : "+r"(a.x), "+r"(a.y), "+r"(a.z), "+r"(a.w), "+r"(b.x), "+r"(b.y), "+r"(b.z), "+r"(b.w));
  }
  __device__ __forceinline__ void fma_into(float (&acc)[8], float w) const {
    acc[0] = fmaf(__uint_as_float(a.x), w, acc[0]); acc[1] = fmaf(__uint_as_float(a.y), w, acc[1]);
    acc[2] = fmaf(__uint_as_float(a.z), w, acc[2]); acc[3] = fmaf(__uint_as_float(a.w), w, acc[3]);
    acc[4] = fmaf(__uint_as_float(b.x), w, acc[4]); acc[5] = fmaf(__uint_as_float(b.y), w, acc[5]);
    acc[6] = fmaf(__uint_as_float(b.z), w, acc[6]); acc[7] = fmaf(__uint_as_float(b.w), w, acc[7]);
  }
};

// Store 8 accumulated channels of one lane.
template <bool kBf16In, bool kBf16Out>
__device__ __forceinline__ void store_lane(void* out, size_t row, int C, int lane, int chunk, const float (&acc)[8]) {
  if (kBf16In) {           // lane owns channels [8*lane, 8*lane+8)
    if (kBf16Out) {
      uint4 o;
      o.x = pack_bf16(acc[0], acc[1]); o.y = pack_bf16(acc[2], acc[3]);
      o.z = pack_bf16(acc[4], acc[5]); o.w = pack_bf16(acc[6], acc[7]);
      *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(out) + row * C + chunk * 256 + lane * 8) = o;
    } else {
      float4* p = reinterpret_cast<float4*>(static_cast<float*>(out) + row * C + chunk * 256 + lane * 8);
      p[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      p[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  } else {                 // lane owns channels [4*lane, +4) and [128+4*lane, +4)
    if (kBf16Out) {
      __nv_bfloat16* p = static_cast<__nv_bfloat16*>(out) + row * C + chunk * 256 + lane * 4;
      uint2 lo, hi;
      lo.x = pack_bf16(acc[0], acc[1]); lo.y = pack_bf16(acc[2], acc[3]);
      hi.x = pack_bf16(acc[4], acc[5]); hi.y = pack_bf16(acc[6], acc[7]);
      *reinterpret_cast<uint2*>(p) = lo;
      *reinterpret_cast<uint2*>(p + 128) = hi;
    } else {
      float* p = static_cast<float*>(out) + row * C + chunk * 256 + lane * 4;
      *reinterpret_cast<float4*>(p) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(p + 128) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}

template <bool kBf16In, bool kBf16Out, int kLevels>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)   // minBlocks=4: ptxas then batches all 16 loads (checked in SASS)
sample_kernel(const SampleParams p) {
  constexpr int kGroup = kBf16In ? kLevels : kLevels / 2;
  __shared__ float s_mat[TC_MAX_CAMS * 16];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < p.N * 16; i += blockDim.x) s_mat[i] = p.lidar2img[(size_t)b * p.N * 16 + i];
  __syncthreads();

  const int q = blockIdx.x * kWarpsPerBlock + warp;
  if (q >= p.Q) return;
  const size_t row = (size_t)b * p.Q + q;

  // -- projection: lane c handles camera c (T:389-409), fp32, reference op order ------------------
  const float rx = p.ref[row * 3 + 0], ry = p.ref[row * 3 + 1], rz = p.ref[row * 3 + 2];
  const float px = __fadd_rn(__fmul_rn(rx, p.pc[3] - p.pc[0]), p.pc[0]);
  const float py = __fadd_rn(__fmul_rn(ry, p.pc[4] - p.pc[1]), p.pc[1]);
  const float pz = __fadd_rn(__fmul_rn(rz, p.pc[5] - p.pc[2]), p.pc[2]);
  float gx = 0.f, gy = 0.f;
  bool valid = false;
  if (lane < p.N) {
    const float* m = s_mat + lane * 16;
    // 4x4 . (px,py,pz,1) in the evaluation order of the batched K=4 SGEMM that torch.matmul dispatches to on
    // sm_100 (found by tools/probe_matmul4.py: bit-identical on 21600/21600 outputs):
    //   (m0*px (+) m1*py)  +  (m2*pz (+) m3*1)   with (+) = FMA
    float cx = __fadd_rn(__fmaf_rn(m[1], py, __fmul_rn(m[0], px)), __fmaf_rn(m[3], 1.0f, __fmul_rn(m[2], pz)));
    float cy = __fadd_rn(__fmaf_rn(m[5], py, __fmul_rn(m[4], px)), __fmaf_rn(m[7], 1.0f, __fmul_rn(m[6], pz)));
    float cz = __fadd_rn(__fmaf_rn(m[9], py, __fmul_rn(m[8], px)), __fmaf_rn(m[11], 1.0f, __fmul_rn(m[10], pz)));
    const float eps = 1e-5f;
    valid = cz > eps;
    const float zc = fmaxf(cz, eps);
    // `x /= python_scalar` on a CUDA tensor is x * (1/scalar) in ATen (BinaryDivTrueKernel: cpu-scalar fast path),
    // not a true division; tensor / tensor (the perspective divide) is IEEE division.
    float u = __fmul_rn(__fdiv_rn(cx, zc), p.inv_w);
    float v = __fmul_rn(__fdiv_rn(cy, zc), p.inv_h);
    gx = __fmul_rn(__fadd_rn(u, -0.5f), 2.0f);
    gy = __fmul_rn(__fadd_rn(v, -0.5f), 2.0f);
    valid = valid && (gx > -1.0f) && (gx < 1.0f) && (gy > -1.0f) && (gy < 1.0f);
    if (p.mask) p.mask[row * p.N + lane] = valid ? 1 : 0;
  }
  unsigned vset = __ballot_sync(0xffffffffu, valid);

  // -- sigmoid(attention logits): lane i < N*L holds weight i = cam*L + level ----------------------
  float wgt = 0.f;
  if (lane < p.N * kLevels) wgt = sigmoid_f32(p.logits[row * (size_t)(p.N * kLevels) + lane]);

  const int chunks = p.C >> 8;
  for (int chunk = 0; chunk < chunks; ++chunk) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;

    unsigned todo = vset;
    while (todo) {
      const int cam = __ffs(todo) - 1;
      todo &= todo - 1;
      const float x = __shfl_sync(0xffffffffu, gx, cam);
      const float y = __shfl_sync(0xffffffffu, gy, cam);

      // kGroup levels at a time: issue all 4*kGroup texel loads (16 independent 128-bit requests per
      // lane for bf16, 2 levels x 8 for fp32), fence, then consume.
#pragma unroll
      for (int l0 = 0; l0 < kLevels; l0 += kGroup) {
        Texel<kBf16In> t[kGroup][4];
        float cw[kGroup][4];
        size_t addr[kGroup][4];
        // phase 1: coordinates, corner weights and texel indices of the whole group (no loads yet)
#pragma unroll
        for (int g = 0; g < kGroup; ++g) {
          const int l = l0 + g;
          const int H = p.H[l], W = p.W[l];
          const float ix = unnormalize(x, W), iy = unnormalize(y, H);
          const float fx0 = floorf(ix), fy0 = floorf(iy);
          const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
          const float wx1 = ix - fx0, wy1 = iy - fy0;                       // (ix - ix_nw), (iy - iy_nw)
          const float wx0 = (fx0 + 1.0f) - ix, wy0 = (fy0 + 1.0f) - iy;     // (ix_se - ix), (iy_se - iy)
          const float lw = __shfl_sync(0xffffffffu, wgt, cam * kLevels + l);
          const bool vx0 = (x0 >= 0) & (x0 < W), vx1 = (x1 >= 0) & (x1 < W);
          const bool vy0 = (y0 >= 0) & (y0 < H), vy1 = (y1 >= 0) & (y1 < H);
          // zeros padding: out-of-range corners get weight 0 and a clamped (always legal) address
          cw[g][0] = (vx0 & vy0) ? wx0 * wy0 * lw : 0.f;   // nw
          cw[g][1] = (vx1 & vy0) ? wx1 * wy0 * lw : 0.f;   // ne
          cw[g][2] = (vx0 & vy1) ? wx0 * wy1 * lw : 0.f;   // sw
          cw[g][3] = (vx1 & vy1) ? wx1 * wy1 * lw : 0.f;   // se
          const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x1, 0), W - 1);
          const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y1, 0), H - 1);
          const size_t plane = ((size_t)b * p.N + cam) * H;
          addr[g][0] = (plane + cy0) * W + cx0;
          addr[g][1] = (plane + cy0) * W + cx1;
          addr[g][2] = (plane + cy1) * W + cx0;
          addr[g][3] = (plane + cy1) * W + cx1;
        }
        // phase 2: all loads back to back
#pragma unroll
        for (int g = 0; g < kGroup; ++g)
#pragma unroll
          for (int c = 0; c < 4; ++c) t[g][c].load(p.feat[l0 + g], addr[g][c], p.C, lane, chunk);
#pragma unroll
        for (int g = 0; g < kGroup; ++g)
#pragma unroll
          for (int c = 0; c < 4; ++c) t[g][c].pin();
#pragma unroll
        for (int g = 0; g < kGroup; ++g)
#pragma unroll
          for (int c = 0; c < 4; ++c) t[g][c].fma_into(acc, cw[g][c]);
      }
    }
    store_lane<kBf16In, kBf16Out>(p.out, row, p.C, lane, chunk, acc);
  }
}

// ---- NCHW fp32 -> NHWC (fp32|bf16) tiled transpose ------------------------------------------------
template <bool kBf16Out>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ src, void* __restrict__ dst,
                                                            int C, int HW) {
  __shared__ float tile[32][33];
  const size_t plane = blockIdx.z;
  const int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    int c = c0 + i, hw = hw0 + tx;
    tile[i][tx] = (c < C && hw < HW) ? src[(plane * C + c) * (size_t)HW + hw] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int hw = hw0 + i, c = c0 + tx;
    if (hw < HW && c < C) {
      float v = tile[tx][i];
      size_t o = (plane * HW + hw) * (size_t)C + c;
      if (kBf16Out) static_cast<__nv_bfloat16*>(dst)[o] = __float2bfloat16_rn(v);
      else static_cast<float*>(dst)[o] = v;
    }
  }
}

}  // namespace
}  // namespace tc

extern "C" int tc_sample_fwd(const tc_sample_args* a, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(a != nullptr, TC_ERR_NULL, "tc_sample_fwd: args is NULL");
  if (a->B == 0 || a->Q == 0) return TC_OK;         // empty batch / query set: nothing to do
  TC_REQUIRE(a->ref && a->lidar2img && a->attn_logits && a->out, TC_ERR_NULL, "tc_sample_fwd: NULL tensor pointer");
  TC_REQUIRE(a->num_levels == 4, TC_ERR_SHAPE, "tc_sample_fwd: num_levels must be 4 (got %d)", a->num_levels);
  TC_REQUIRE(a->N >= 1 && a->N <= TC_MAX_CAMS && a->N * a->num_levels <= 32, TC_ERR_SHAPE,
             "tc_sample_fwd: num cams %d unsupported", a->N);
  TC_REQUIRE(a->C > 0 && a->C % 256 == 0, TC_ERR_SHAPE, "tc_sample_fwd: C must be a multiple of 256 (got %d)", a->C);
  TC_REQUIRE(a->B >= 0 && a->Q >= 0 && a->B <= 65535, TC_ERR_SHAPE, "tc_sample_fwd: bad B/Q");
  TC_REQUIRE(a->feat_dtype == TC_F32 || a->feat_dtype == TC_BF16, TC_ERR_DTYPE, "tc_sample_fwd: bad feat dtype");
  TC_REQUIRE(a->out_dtype == TC_F32 || a->out_dtype == TC_BF16, TC_ERR_DTYPE, "tc_sample_fwd: bad out dtype");
  TC_REQUIRE(aligned16(a->out), TC_ERR_ALIGN, "tc_sample_fwd: out must be 16-byte aligned");
  if (a->B == 0 || a->Q == 0) return TC_OK;
  SampleParams p;
  for (int l = 0; l < 4; ++l) {
    TC_REQUIRE(a->feat[l] != nullptr, TC_ERR_NULL, "tc_sample_fwd: feat[%d] is NULL", l);
    TC_REQUIRE(aligned16(a->feat[l]), TC_ERR_ALIGN, "tc_sample_fwd: feat[%d] must be 16-byte aligned", l);
    TC_REQUIRE(a->H[l] > 0 && a->W[l] > 0, TC_ERR_SHAPE, "tc_sample_fwd: level %d has empty extent", l);
    p.feat[l] = a->feat[l]; p.H[l] = a->H[l]; p.W[l] = a->W[l];
  }
  p.B = a->B; p.N = a->N; p.Q = a->Q; p.C = a->C;
  p.ref = a->ref; p.lidar2img = a->lidar2img; p.logits = a->attn_logits;
  for (int i = 0; i < 6; ++i) p.pc[i] = a->pc_range[i];
  p.inv_w = 1.0f / a->img_w; p.inv_h = 1.0f / a->img_h;
  p.out = a->out; p.mask = a->mask;
  dim3 grid((a->Q + kWarpsPerBlock - 1) / kWarpsPerBlock, a->B);
  dim3 block(kWarpsPerBlock * 32);
  cudaStream_t s = as_stream(stream);
  const bool bi = a->feat_dtype == TC_BF16, bo = a->out_dtype == TC_BF16;
  if (bi && bo) sample_kernel<true, true, 4><<<grid, block, 0, s>>>(p);
  else if (bi) sample_kernel<true, false, 4><<<grid, block, 0, s>>>(p);
  else if (bo) sample_kernel<false, true, 4><<<grid, block, 0, s>>>(p);
  else sample_kernel<false, false, 4><<<grid, block, 0, s>>>(p);
  count_launch();
  return check_launch("tc_sample_fwd");
}

extern "C" int tc_nchw_to_nhwc(const float* src, void* dst, int32_t dst_dtype, int32_t planes, int32_t C,
                               int32_t HW, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(src && dst, TC_ERR_NULL, "tc_nchw_to_nhwc: NULL pointer");
  TC_REQUIRE(planes >= 0 && C > 0 && HW > 0 && planes <= 65535, TC_ERR_SHAPE, "tc_nchw_to_nhwc: bad shape");
  TC_REQUIRE(dst_dtype == TC_F32 || dst_dtype == TC_BF16, TC_ERR_DTYPE, "tc_nchw_to_nhwc: bad dtype");
  if (planes == 0) return TC_OK;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, planes);
  if (dst_dtype == TC_BF16) nchw_to_nhwc_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(src, dst, C, HW);
  else nchw_to_nhwc_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(src, dst, C, HW);
  count_launch();
  return check_launch("tc_nchw_to_nhwc");
}
