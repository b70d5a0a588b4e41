// K3 (exact path): Linear + fused epilogue on the CUDA cores, fp32 FFMA accumulation.
// This is the parity mode (fp32 operands, 1e-5 vs the reference) and the fallback for the tiny odd
// shapes of the bf16 build (N = 10, 24; K = 36).  The tensor-core (tcgen05) path lives in linear_tc.cu.
//
// Tile: 32 rows x (32 * CPL) columns per 256-thread block, K step 16, CPL = 8 (256 columns) or 2 (N <= 64: the first
// radar feature layer, 36 -> 64, would leave 24 of 32 lanes idle in the wide tile).  Thread (warp w, lane l) owns rows
// 4w..4w+3 and columns CPL*l..CPL*l+CPL-1, so a full output row lives in one warp and the LayerNorm epilogue is a
// pair of warp reductions - no second pass over HBM.
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace tc {
namespace {

constexpr int BM = 32, BK = 16;

struct LinearParams {
  const void* A; long long lda;
  const void* W; long long ldw;
  int M, N, K;
  const float* bias;
  const float* row_bias; int row_bias_period; long long ld_row_bias;
  const uint8_t* row_gate;
  const float* residual; long long ld_residual;
  const float* residual2; long long ld_residual2;
  const float* ln_gamma; const float* ln_beta; float ln_eps;
  int relu;
  const float* post_add; long long ld_post_add;
  float* out_f32; long long ld_out_f32;
  __nv_bfloat16* out_bf16; long long ld_out_bf16;
  int out16;             // TC_BF16, TC_BF16X2 (hi at column n, lo at column N + n) or TC_F16
  TailParams tail;
  int vec_a, vec_w;      // 1: rows are 16-byte (fp32) / 8-byte (bf16) aligned and K % 4 == 0
};

template <typename T>
__device__ __forceinline__ float4 load4(const T* base, long long ld, int row, int rows, int k, int K, int vec);

template <>
__device__ __forceinline__ float4 load4<float>(const float* base, long long ld, int row, int rows, int k, int K, int vec) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row >= rows) return r;
  const float* p = base + (long long)row * ld + k;
  if (vec && k + 3 < K) return *reinterpret_cast<const float4*>(p);
  if (k + 0 < K) r.x = p[0];
  if (k + 1 < K) r.y = p[1];
  if (k + 2 < K) r.z = p[2];
  if (k + 3 < K) r.w = p[3];
  return r;
}

template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* base, long long ld, int row, int rows,
                                                       int k, int K, int vec) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row >= rows) return r;
  const __nv_bfloat16* p = base + (long long)row * ld + k;
  if (vec && k + 3 < K) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    return make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
  }
  if (k + 0 < K) r.x = __bfloat162float(p[0]);
  if (k + 1 < K) r.y = __bfloat162float(p[1]);
  if (k + 2 < K) r.z = __bfloat162float(p[2]);
  if (k + 3 < K) r.w = __bfloat162float(p[3]);
  return r;
}

// split bf16 operand ([rows, 2K]: hi | lo): the value is hi + lo
struct SplitBf16 {};
template <>
__device__ __forceinline__ float4 load4<SplitBf16>(const SplitBf16* base, long long ld, int row, int rows, int k, int K, int vec) {
  const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(base);
  const float4 hi = load4<__nv_bfloat16>(b, ld, row, rows, k, K, vec);
  const float4 lo = load4<__nv_bfloat16>(b + K, ld, row, rows, k, K, vec);
  return make_float4(hi.x + lo.x, hi.y + lo.y, hi.z + lo.z, hi.w + lo.w);
}

template <typename TA, typename TW, int CPL>
__global__ void __launch_bounds__(256) linear_simt_kernel(const LinearParams p) {
  constexpr int BN = 32 * CPL;
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Ws[BK][BN];
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const TA* A = static_cast<const TA*>(p.A);
  const TW* W = static_cast<const TW*>(p.W);

  float acc[4][CPL];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < CPL; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    if (tid < 128) {                        // A tile: 32 rows x 16 k
      const int r = tid >> 2, kq = (tid & 3) * 4;
      float4 v = load4<TA>(A, p.lda, m0 + r, p.M, k0 + kq, p.K, p.vec_a);
      As[kq + 0][r] = v.x; As[kq + 1][r] = v.y; As[kq + 2][r] = v.z; As[kq + 3][r] = v.w;
    }
    if (tid < BN) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {         // W tile: BN rows x 16 k
        const int kq = i * 4;
        float4 v = load4<TW>(W, p.ldw, n0 + tid, p.N, k0 + kq, p.K, p.vec_w);
        Ws[kq + 0][tid] = v.x; Ws[kq + 1][tid] = v.y; Ws[kq + 2][tid] = v.z; Ws[kq + 3][tid] = v.w;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][warp * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      float wv[CPL];
      if (CPL == 8) {
        const float4 w0 = *reinterpret_cast<const float4*>(&Ws[k][lane * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&Ws[k][lane * 8 + 4]);
        wv[0] = w0.x; wv[1] = w0.y; wv[2 % CPL] = w0.z; wv[3 % CPL] = w0.w;
        wv[4 % CPL] = w1.x; wv[5 % CPL] = w1.y; wv[6 % CPL] = w1.z; wv[7 % CPL] = w1.w;
      } else {
        const float2 w0 = *reinterpret_cast<const float2*>(&Ws[k][lane * 2]);
        wv[0] = w0.x; wv[1] = w0.y;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CPL; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue ---------------------------------------------------------------------------------
  const int nbase = n0 + lane * CPL;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + warp * 4 + i;
    if (m >= p.M) continue;                 // warp-uniform
    float y[CPL];
    const bool gate = p.row_gate ? (p.row_gate[m] != 0) : true;
    const float* rb = p.row_bias ? p.row_bias + (long long)(m % p.row_bias_period) * p.ld_row_bias : nullptr;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int n = nbase + j;
      float v = 0.f;
      if (n < p.N) {
        v = acc[i][j];
        if (p.bias) v += p.bias[n];
        if (rb) v += rb[n];
        if (!gate) v = 0.f;
        if (p.residual) v += p.residual[(long long)m * p.ld_residual + n];
        if (p.residual2) v += p.residual2[(long long)m * p.ld_residual2 + n];
      }
      y[j] = v;
    }
    if (p.ln_gamma) {                       // whole row is inside this warp (host guarantees N <= 256)
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) s += (nbase + j < p.N) ? y[j] : 0.f;
      const float mean = warp_sum(s) / (float)p.N;
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const float d = (nbase + j < p.N) ? y[j] - mean : 0.f;
        sq += d * d;
      }
      const float rstd = rsqrtf(warp_sum(sq) / (float)p.N + p.ln_eps);
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int n = nbase + j;
        if (n < p.N) y[j] = (y[j] - mean) * rstd * p.ln_gamma[n] + p.ln_beta[n];
      }
    }
    if (p.tail.kind != TC_TAIL_NONE && n0 == 0) {
      // columns 0..7 of the row live in lanes 0 .. 8/CPL-1: gather them, run the tail in lane 0, scatter the updates back
      float r8[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float mine = p.relu ? fmaxf(y[c % CPL], 0.f) : y[c % CPL];
        r8[c] = __shfl_sync(0xffffffffu, mine, c / CPL);
      }
      if (lane == 0) apply_tail(p.tail, m, r8);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float upd = __shfl_sync(0xffffffffu, r8[c], 0);
        if (lane == c / CPL) y[c % CPL] = upd;          // relu (if any) already applied; values >= 0 stay put below
      }
    }
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int n = nbase + j;
      if (n >= p.N) continue;
      float v = (p.relu && p.tail.kind == TC_TAIL_NONE) ? fmaxf(y[j], 0.f) : y[j];
      if (p.post_add) v += p.post_add[(long long)m * p.ld_post_add + n];
      if (p.out_f32) p.out_f32[(long long)m * p.ld_out_f32 + n] = v;
      if (p.out_bf16) {
        __nv_bfloat16* d16 = p.out_bf16 + (long long)m * p.ld_out_bf16 + n;
        if (p.out16 == TC_F16) {
          *reinterpret_cast<__half*>(d16) = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
        } else {
          const __nv_bfloat16 hi = __float2bfloat16_rn(v);
          *d16 = hi;
          if (p.out16 == TC_BF16X2) d16[p.N] = __float2bfloat16_rn(v - __bfloat162float(hi));
        }
      }
    }
  }
}

// ---- fused 3 -> C point embedding: ReLU(LN(Linear3(f(x)))) ---------------------------------------
struct PointEmbedParams {
  const float* x; long long ldx; int M, C, logit;
  const float* weight; const float* bias; const float* gamma; const float* beta; float eps;
  float* out_f32; __nv_bfloat16* out_bf16; int split;
};

constexpr int kPeMaxPerLane = 32;   // C <= 1024

__global__ void __launch_bounds__(128) point_embed_kernel(const PointEmbedParams p) {
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (m >= p.M) return;
  float x0 = p.x[(long long)m * p.ldx + 0], x1 = p.x[(long long)m * p.ldx + 1], x2 = p.x[(long long)m * p.ldx + 2];
  if (p.logit) { x0 = logit_f32(x0); x1 = logit_f32(x1); x2 = logit_f32(x2); }
  const int per = p.C >> 5;
  float y[kPeMaxPerLane];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kPeMaxPerLane; ++j) {
    if (j < per) {
      const int c = j * 32 + lane;
      // x . w in k order, then + bias (addmm: bias + x W^T; K = 3)
      float v = fmaf(x2, p.weight[c * 3 + 2], fmaf(x1, p.weight[c * 3 + 1], x0 * p.weight[c * 3 + 0])) + p.bias[c];
      y[j] = v;
      s += v;
    }
  }
  const float mean = warp_sum(s) / (float)p.C;
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < kPeMaxPerLane; ++j)
    if (j < per) { const float d = y[j] - mean; sq += d * d; }
  const float rstd = rsqrtf(warp_sum(sq) / (float)p.C + p.eps);
#pragma unroll
  for (int j = 0; j < kPeMaxPerLane; ++j) {
    if (j < per) {
      const int c = j * 32 + lane;
      const float v = fmaxf((y[j] - mean) * rstd * p.gamma[c] + p.beta[c], 0.f);
      if (p.out_f32) p.out_f32[(long long)m * p.C + c] = v;
      if (p.out_bf16) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        if (p.split) {
          p.out_bf16[(long long)m * 2 * p.C + c] = hi;
          p.out_bf16[(long long)m * 2 * p.C + p.C + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
        } else {
          p.out_bf16[(long long)m * p.C + c] = hi;
        }
      }
    }
  }
}

// C == 256 specialisation: persistent warps, lane owns channels [8*lane, 8*lane+8) and keeps their 24 weights,
// bias, gain and shift in registers across rows.  The parameters reach the registers through one coalesced,
// transposed shared-memory copy per CTA (a direct per-lane read of weight[c][k] costs 24 L1 wavefronts per load).
__global__ void __launch_bounds__(256) point_embed256_kernel(const PointEmbedParams p) {
  __shared__ __align__(16) float s_par[6 * 256];          // w[k=0..2][c], bias[c], gamma[c], beta[c]
  for (int f = threadIdx.x; f < 768; f += 256) s_par[(f % 3) * 256 + f / 3] = __ldg(p.weight + f);
  s_par[768 + threadIdx.x] = __ldg(p.bias + threadIdx.x);
  s_par[1024 + threadIdx.x] = __ldg(p.gamma + threadIdx.x);
  s_par[1280 + threadIdx.x] = __ldg(p.beta + threadIdx.x);
  __syncthreads();
  pdl_trigger();
  pdl_wait();            // parameters above are constants; x comes from the previous kernel
  const int lane = threadIdx.x & 31;
  const int c0 = lane * 8;
  float w[8][3], bias[8], gamma[8], beta[8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float4 t = *reinterpret_cast<const float4*>(s_par + k * 256 + c0 + 4 * h);
      w[4 * h + 0][k] = t.x; w[4 * h + 1][k] = t.y; w[4 * h + 2][k] = t.z; w[4 * h + 3][k] = t.w;
    }
    const float4 tb = *reinterpret_cast<const float4*>(s_par + 768 + c0 + 4 * h);
    const float4 tg = *reinterpret_cast<const float4*>(s_par + 1024 + c0 + 4 * h);
    const float4 te = *reinterpret_cast<const float4*>(s_par + 1280 + c0 + 4 * h);
    bias[4 * h + 0] = tb.x; bias[4 * h + 1] = tb.y; bias[4 * h + 2] = tb.z; bias[4 * h + 3] = tb.w;
    gamma[4 * h + 0] = tg.x; gamma[4 * h + 1] = tg.y; gamma[4 * h + 2] = tg.z; gamma[4 * h + 3] = tg.w;
    beta[4 * h + 0] = te.x; beta[4 * h + 1] = te.y; beta[4 * h + 2] = te.z; beta[4 * h + 3] = te.w;
  }
  const int stride = gridDim.x * 8;
  for (int m = blockIdx.x * 8 + (threadIdx.x >> 5); m < p.M; m += stride) {
    const float* xr = p.x + (long long)m * p.ldx;
    float x0 = __ldg(xr), x1 = __ldg(xr + 1), x2 = __ldg(xr + 2);
    if (p.logit) { x0 = logit_f32(x0); x1 = logit_f32(x1); x2 = logit_f32(x2); }
    float y[8], s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      y[i] = fmaf(x2, w[i][2], fmaf(x1, w[i][1], x0 * w[i][0])) + bias[i];     // k order, then + bias (addmm)
      s += y[i];
    }
    const float mean = warp_sum(s) * (1.0f / 256.0f);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = y[i] - mean; sq = fmaf(d, d, sq); }
    const float rstd = rsqrtf(warp_sum(sq) * (1.0f / 256.0f) + p.eps);
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = fmaxf((y[i] - mean) * rstd * gamma[i] + beta[i], 0.f);
    if (p.out_f32) {
      float4* o = reinterpret_cast<float4*>(p.out_f32 + (long long)m * 256 + c0);
      o[0] = make_float4(y[0], y[1], y[2], y[3]);
      o[1] = make_float4(y[4], y[5], y[6], y[7]);
    }
    if (p.out_bf16) {
      uint4 u;
      u.x = pack_bf16(y[0], y[1]); u.y = pack_bf16(y[2], y[3]); u.z = pack_bf16(y[4], y[5]); u.w = pack_bf16(y[6], y[7]);
      if (p.split) {              // [M, 512] = hi | lo
        *reinterpret_cast<uint4*>(p.out_bf16 + (long long)m * 512 + c0) = u;
        uint4 l;
        l.x = pack_bf16(y[0] - bf16_lo(u.x), y[1] - bf16_hi(u.x)); l.y = pack_bf16(y[2] - bf16_lo(u.y), y[3] - bf16_hi(u.y));
        l.z = pack_bf16(y[4] - bf16_lo(u.z), y[5] - bf16_hi(u.z)); l.w = pack_bf16(y[6] - bf16_lo(u.w), y[7] - bf16_hi(u.w));
        *reinterpret_cast<uint4*>(p.out_bf16 + (long long)m * 512 + 256 + c0) = l;
      } else {
        *reinterpret_cast<uint4*>(p.out_bf16 + (long long)m * 256 + c0) = u;
      }
    }
  }
}

}  // namespace

int linear_simt_launch(const tc_linear_args* a, cudaStream_t s) {
  LinearParams p;
  p.A = a->A; p.lda = a->lda; p.W = a->W; p.ldw = a->ldw;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.bias = a->bias;
  p.row_bias = a->row_bias; p.row_bias_period = a->row_bias_period > 0 ? a->row_bias_period : 1;
  p.ld_row_bias = a->ld_row_bias;
  p.row_gate = a->row_gate;
  p.residual = a->residual; p.ld_residual = a->ld_residual;
  p.residual2 = a->residual2; p.ld_residual2 = a->ld_residual2;
  p.ln_gamma = a->ln_gamma; p.ln_beta = a->ln_beta; p.ln_eps = a->ln_eps;
  p.relu = a->relu;
  p.post_add = a->post_add; p.ld_post_add = a->ld_post_add;
  p.out_f32 = a->out_f32; p.ld_out_f32 = a->ld_out_f32;
  p.out_bf16 = static_cast<__nv_bfloat16*>(a->out_bf16); p.ld_out_bf16 = a->ld_out_bf16;
  p.out16 = a->out16_dtype == 0 ? TC_BF16 : a->out16_dtype;
  p.tail = make_tail(a);
  const int ea = a->a_dtype == TC_F32 ? 4 : 2, ew = a->w_dtype == TC_F32 ? 4 : 2;
  p.vec_a = (a->K % 4 == 0) && ((a->lda * ea) % (4 * ea) == 0) && ((reinterpret_cast<uintptr_t>(a->A) % (4 * ea)) == 0);
  p.vec_w = (a->K % 4 == 0) && ((a->ldw * ew) % (4 * ew) == 0) && ((reinterpret_cast<uintptr_t>(a->W) % (4 * ew)) == 0);
  // N <= 64 without LayerNorm (whose row must sit in one warp: any N <= 256 does in the wide tile, N <= 64 in the narrow one)
  const bool narrow = a->N <= 64;
  const int bn = narrow ? 64 : 256;
  dim3 grid((a->M + BM - 1) / BM, (a->N + bn - 1) / bn);
#define TC_SIMT_LAUNCH(TA, TW)                                                                          \
  (narrow ? launch(linear_simt_kernel<TA, TW, 2>, grid, dim3(256), 0, s, 1u, p)                         \
          : launch(linear_simt_kernel<TA, TW, 8>, grid, dim3(256), 0, s, 1u, p))
  if (a->a_dtype == TC_BF16X2 && a->w_dtype == TC_BF16X2) TC_SIMT_LAUNCH(SplitBf16, SplitBf16);
  else if (a->a_dtype == TC_F32 && a->w_dtype == TC_F32) TC_SIMT_LAUNCH(float, float);
  else if (a->a_dtype == TC_BF16 && a->w_dtype == TC_BF16) TC_SIMT_LAUNCH(__nv_bfloat16, __nv_bfloat16);
  else if (a->a_dtype == TC_F32 && a->w_dtype == TC_BF16) TC_SIMT_LAUNCH(float, __nv_bfloat16);
  else TC_SIMT_LAUNCH(__nv_bfloat16, float);
#undef TC_SIMT_LAUNCH
  count_launch();
  return check_launch("tc_linear(simt)");
}

}  // namespace tc

extern "C" int tc_point_embed(const tc_point_embed_args* a, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(a != nullptr, TC_ERR_NULL, "tc_point_embed: args is NULL");
  TC_REQUIRE(a->x && a->weight && a->bias && a->ln_gamma && a->ln_beta, TC_ERR_NULL, "tc_point_embed: NULL pointer");
  TC_REQUIRE(a->out_f32 || a->out_bf16, TC_ERR_NULL, "tc_point_embed: no output");
  TC_REQUIRE(a->C > 0 && a->C % 32 == 0 && a->C <= 32 * kPeMaxPerLane, TC_ERR_SHAPE, "tc_point_embed: bad C %d", a->C);
  TC_REQUIRE(a->M >= 0 && a->ldx >= 3, TC_ERR_SHAPE, "tc_point_embed: bad M/ldx");
  if (a->M == 0) return TC_OK;
  PointEmbedParams p{a->x, a->ldx, a->M, a->C, a->logit_input, a->weight, a->bias, a->ln_gamma, a->ln_beta,
                     a->ln_eps, a->out_f32, static_cast<__nv_bfloat16*>(a->out_bf16), a->out16_dtype == TC_BF16X2 ? 1 : 0};
  const bool al = (!a->out_f32 || aligned16(a->out_f32)) && (!a->out_bf16 || aligned16(a->out_bf16));
  if (a->C == 256 && al) {
    const int ctas = (a->M + 7) / 8;
    launch(point_embed256_kernel, dim3(ctas < 296 ? ctas : 296), dim3(256), 0, as_stream(stream), 1u, p);      // 2 CTAs per SM x 148
  } else {
    point_embed_kernel<<<(a->M + 3) / 4, 128, 0, as_stream(stream)>>>(p);
  }
  count_launch();
  return check_launch("tc_point_embed");
}
