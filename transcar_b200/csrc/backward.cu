// Backward-pass building blocks of the radar fusion head (the part of TransCAR that trains: tools/train.py:238-252
// freezes the backbone, neck, DETR3D transformer, cls/reg branches and query embedding).  The GEMMs of the backward
// pass are tc_linear calls on transposed operands (dX = dY W, dW = dY^T X); this file holds everything else:
//   tc_transpose        [R,C] -> [C,R] with optional fp32 -> bf16 conversion (operands of the wgrad GEMMs, W^T)
//   tc_colsum           out[n] += sum_m x[m,n]                               (bias gradients)
//   tc_layernorm_fwd    y = LN(x) * gamma + beta, saves mean / rstd          (training forward keeps LN un-fused)
//   tc_layernorm_bwd    dx (+= add), dgamma += , dbeta +=                    (nn.LayerNorm backward, biased variance)
//   tc_mask_grad        dz = dy * (y > 0) * gate[row]                        (ReLU backward and / or the row gate of quirk Q6)
//   tc_attention_sparse_bwd   dq, dk +=, dv += of the masked radar attention (detr3d_head.py:578), same key scan and
//                       bit-exact mask as attention_sparse.cu; the softmax is recomputed, never stored
// All reductions into parameters / shared rows use fp32 atomics (the caller zeroes the gradient bucket once per step).
#include "tc_common.cuh"

namespace tc {
namespace {

// ---- transpose ---------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) transpose_kernel(const TI* __restrict__ src, long long ld_src, TO* __restrict__ dst,
                                                        long long ld_dst, int rows, int cols) {
  __shared__ float tile[32][33];
  pdl_trigger();
  pdl_wait();
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? static_cast<float>(src[(long long)r * ld_src + c]) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;                         // dst[c][r]
    if (c < cols && r < rows) dst[(long long)c * ld_dst + r] = static_cast<TO>(tile[tx][i]);
  }
}

// [R,C] fp32 -> split bf16 [C, 2*Rp] (TC_BF16X2: hi | lo halves of Rp columns each); rows R..Rp-1 are written as zeros so
// that Rp can be the 64-aligned reduction length of a bf16x3 wgrad GEMM (dW = dY^T X reduces over the M rows).
__global__ void __launch_bounds__(256) transpose_split_kernel(const float* __restrict__ src, long long ld_src,
                                                              __nv_bfloat16* __restrict__ dst, long long ld_dst, int rows,
                                                              int rows_pad, int cols) {
  __shared__ float tile[32][33];
  pdl_trigger();
  pdl_wait();
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? src[(long long)r * ld_src + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < cols && r < rows_pad) {
      const float v = tile[tx][i];
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      dst[(long long)c * ld_dst + r] = hi;
      dst[(long long)c * ld_dst + rows_pad + r] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
  }
}

// ---- column sum --------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long long ldx, int M, int N, float* __restrict__ out) {
  __shared__ float part[8][33];
  pdl_trigger();
  pdl_wait();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (n < N)
    for (int m = blockIdx.y * 8 + ty; m < M; m += gridDim.y * 8) s += static_cast<float>(x[(long long)m * ldx + n]);
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
#pragma unroll
    for (int i = 1; i < 8; ++i) s += part[i][tx];
    atomicAdd(out + n, s);
  }
}

// ---- LayerNorm ---------------------------------------------------------------------------------------------
constexpr int kLnMaxPerLane = 32;      // N <= 1024

__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const tc_layernorm_args a) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= a.M) return;
  const float* x = a.x + (long long)m * a.ldx;
  const int per = (a.N + 31) >> 5;
  float v[kLnMaxPerLane];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kLnMaxPerLane; ++j) {
    v[j] = 0.f;
    if (j < per) {
      const int n = j * 32 + lane;
      if (n < a.N) { v[j] = x[n]; s += v[j]; }
    }
  }
  const float mean = warp_sum(s) / (float)a.N;
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < kLnMaxPerLane; ++j)
    if (j < per && j * 32 + lane < a.N) { const float d = v[j] - mean; sq += d * d; }
  const float rstd = rsqrtf(warp_sum(sq) / (float)a.N + a.eps);
  if (lane == 0) {
    if (a.mean) a.mean[m] = mean;
    if (a.rstd) a.rstd[m] = rstd;
  }
#pragma unroll
  for (int j = 0; j < kLnMaxPerLane; ++j) {
    if (j < per) {
      const int n = j * 32 + lane;
      if (n < a.N) {
        float y = (v[j] - mean) * rstd * a.gamma[n] + a.beta[n];
        if (a.relu) y = fmaxf(y, 0.f);
        if (a.y_f32) a.y_f32[(long long)m * a.ldy + n] = y;
        if (a.y_bf16) static_cast<__nv_bfloat16*>(a.y_bf16)[(long long)m * a.ldy + n] = __float2bfloat16_rn(y);
      }
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;  dgamma += sum_m dy * xhat;  dbeta += sum_m dy
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const tc_layernorm_bwd_args a) {
  extern __shared__ float sh[];                     // [2][N] per-CTA partial dgamma / dbeta
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 2 * a.N; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int per = (a.N + 31) >> 5;
  for (int m = blockIdx.x * 8 + warp; m < a.M; m += gridDim.x * 8) {
    const float* x = a.x + (long long)m * a.ldx;
    const float* dy = a.dy + (long long)m * a.ld_dy;
    const float mean = a.mean[m], rstd = a.rstd[m];
    float xh[kLnMaxPerLane], g[kLnMaxPerLane];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxPerLane; ++j) {
      xh[j] = 0.f; g[j] = 0.f;
      if (j < per) {
        const int n = j * 32 + lane;
        if (n < a.N) {
          const float d = dy[n];
          xh[j] = (x[n] - mean) * rstd;
          g[j] = d * a.gamma[n];
          s1 += g[j];
          s2 = fmaf(g[j], xh[j], s2);
          atomicAdd(&sh[n], d * xh[j]);              // shared-memory atomics: 8 warps of the CTA hit distinct rows
          atomicAdd(&sh[a.N + n], d);
        }
      }
    }
    const float m1 = warp_sum(s1) / (float)a.N, m2 = warp_sum(s2) / (float)a.N;
#pragma unroll
    for (int j = 0; j < kLnMaxPerLane; ++j) {
      if (j < per) {
        const int n = j * 32 + lane;
        if (n < a.N) {
          float v = rstd * (g[j] - m1 - xh[j] * m2);
          if (a.add) v += a.add[(long long)m * a.ld_add + n];
          a.dx[(long long)m * a.ld_dx + n] = v;
        }
      }
    }
  }
  __syncthreads();
  for (int n = threadIdx.x; n < a.N; n += blockDim.x) {
    if (a.dgamma) atomicAdd(a.dgamma + n, sh[n]);
    if (a.dbeta) atomicAdd(a.dbeta + n, sh[a.N + n]);
  }
}

// ---- ReLU backward / row gate ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mask_grad_kernel(const float* __restrict__ dy, long long ld_dy, const float* __restrict__ y,
                                                        long long ld_y, const uint8_t* __restrict__ gate, float* __restrict__ dz,
                                                        long long ld_dz, int M, int N) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int m = (int)(i / N), n = (int)(i % N);
  float v = dy[(long long)m * ld_dy + n];
  if (y && !(y[(long long)m * ld_y + n] > 0.f)) v = 0.f;
  if (gate && gate[m] == 0) v = 0.f;
  dz[(long long)m * ld_dz + n] = v;
}

// ---- masked radar attention backward ---------------------------------------------------------------------------------
constexpr int kWarps = 8;
constexpr int kKeyTile = 2048;

struct SparseBwdParams {
  const float* q; const float* k; const float* v; const float* dout;
  long long ldq, ldk, ldv, ldo, qbs, kbs, vbs;
  int B, Lq, Lk;
  float scale;
  const float* geom; const float* key_xy;
  float* dq; long long ld_dq;
  float* dk; float* dv; long long ld_dk, ld_dv, dkbs, dvbs;
  DropoutRng rng;
};

__device__ __forceinline__ void ld8(const float* p, float (&d)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}

// One warp per (sample, query); lane owns 8 channels, a head = 4 lanes (see attention_sparse.cu).  Two passes over the
// allowed keys: (1) row max and normaliser per head, (2) p, dP = dO.v, dS = p (dP - dO.O), accumulate dq / dk / dv.
__global__ void __launch_bounds__(kWarps * 32) attention_sparse_bwd_kernel(const SparseBwdParams p) {
  __shared__ float s_kx[kKeyTile], s_ky[kKeyTile], s_kn[kKeyTile];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  const bool active = q < p.Lq;
  const long long row = (long long)b * p.Lq + (active ? q : 0);
  const float* g = p.geom + row * 8;
  const float cx = __ldg(g + 0), cy = __ldg(g + 1), fx = __ldg(g + 2), fy = __ldg(g + 3);
  const float rx = __ldg(g + 4), ry = __ldg(g + 5), radius = __ldg(g + 6);
  const Circle cc = make_circle(cx, cy), cf = make_circle(fx, fy), cr = make_circle(rx, ry);
  const float span = fmaxf(sqrtf((fx - cx) * (fx - cx) + (fy - cy) * (fy - cy)),
                           sqrtf((rx - cx) * (rx - cx) + (ry - cy) * (ry - cy)));
  const float bound = radius + span + 0.25f;
  const float bound2 = bound * bound;

  float qv[8], dO[8], dq[8];
  ld8(p.q + (long long)b * p.qbs + (long long)(active ? q : 0) * p.ldq + lane * 8, qv);
  ld8(p.dout + row * p.ldo + lane * 8, dO);
#pragma unroll
  for (int i = 0; i < 8; ++i) { qv[i] *= p.scale; dq[i] = 0.f; }
  const float2* keys = reinterpret_cast<const float2*>(p.key_xy) + (long long)b * p.Lk;

  float m_run = -INFINITY, l_run = 0.f;                   // per head: running max and normaliser
  float oacc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) oacc[i] = 0.f;

  bool live = active;
  for (int pass = 0; pass < 2; ++pass) {
    float inv_l = 0.f, D = 0.f;
    if (pass == 1 && !(l_run > 0.f)) live = false;          // no allowed key: all gradients of this row are zero
    if (pass == 1 && live) {                                // (the warp keeps taking part in the block barriers below)
      inv_l = 1.0f / l_run;
      // D = dO . O per head, O = oacc / l
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t = fmaf(dO[i], oacc[i] * inv_l, t);
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      D = t;
    }
    for (int t0 = 0; t0 < p.Lk; t0 += kKeyTile) {
      const int nt = min(kKeyTile, p.Lk - t0);
      __syncthreads();
      for (int i = threadIdx.x; i < nt; i += kWarps * 32) {
        const float2 xy = __ldg(keys + t0 + i);
        s_kx[i] = xy.x; s_ky[i] = xy.y; s_kn[i] = key_norm(xy.x, xy.y);
      }
      __syncthreads();
      if (!live) continue;
      for (int k0 = 0; k0 < nt; k0 += 32) {
        const int key = k0 + lane;
        float kx = 0.f, ky = 0.f;
        bool cand = false;
        if (key < nt) {
          kx = s_kx[key]; ky = s_ky[key];
          const float dx = kx - cx, dy = ky - cy;
          cand = !(fmaf(dx, dx, dy * dy) >= bound2);
        }
        if (!__any_sync(0xffffffffu, cand)) continue;
        bool ok = false;
        if (cand) ok = radar_allowed(cc, cf, cr, radius, kx, ky, s_kn[key]);
        unsigned todo = __ballot_sync(0xffffffffu, ok);
        while (todo) {
          const int j = t0 + k0 + __ffs(todo) - 1;
          todo &= todo - 1;
          float kv[8], vv[8];
          ld8(p.k + (long long)b * p.kbs + (long long)j * p.ldk + lane * 8, kv);
          ld8(p.v + (long long)b * p.vbs + (long long)j * p.ldv + lane * 8, vv);
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) s = fmaf(qv[i], kv[i], s);
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          // same dropout decision as the forward kernel: O = sum_j (drop_j p_j) v_j, normaliser un-dropped
          const float drop = p.rng.p > 0.f ? dropout_scale(p.rng, (uint32_t)((b * 8 + (lane >> 2)) * p.Lq + q), (uint32_t)j) : 1.0f;
          if (pass == 0) {
            const float m_new = fmaxf(m_run, s);
            const float corr = expf(m_run - m_new);
            const float pj = expf(s - m_new);
            l_run = fmaf(l_run, corr, pj);
#pragma unroll
            for (int i = 0; i < 8; ++i) oacc[i] = fmaf(oacc[i], corr, pj * drop * vv[i]);
            m_run = m_new;
          } else {
            const float pt = expf(s - m_run) * inv_l;       // softmax probability
            const float pj = pt * drop;                      // what multiplied V in the forward
            float dP = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) dP = fmaf(dO[i], vv[i], dP);
            dP += __shfl_xor_sync(0xffffffffu, dP, 1);
            dP += __shfl_xor_sync(0xffffffffu, dP, 2);
            const float dS = pt * (dP * drop - D);
            float* dkr = p.dk + (long long)b * p.dkbs + (long long)j * p.ld_dk + lane * 8;
            float* dvr = p.dv + (long long)b * p.dvbs + (long long)j * p.ld_dv + lane * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              dq[i] = fmaf(dS, kv[i], dq[i]);
              atomicAdd(dkr + i, dS * qv[i]);              // qv already carries the 1/sqrt(d) scale
              atomicAdd(dvr + i, pj * dO[i]);
            }
          }
        }
      }
    }
  }
  if (!active) return;
  float4* o = reinterpret_cast<float4*>(p.dq + row * p.ld_dq + lane * 8);
  o[0] = make_float4(dq[0] * p.scale, dq[1] * p.scale, dq[2] * p.scale, dq[3] * p.scale);
  o[1] = make_float4(dq[4] * p.scale, dq[5] * p.scale, dq[6] * p.scale, dq[7] * p.scale);
}

}  // namespace
}  // namespace tc

// ------------------------------------------------------------------------------------------------ C ABI
namespace tc {
namespace {

// ---- dense attention backward (decoder self-attention, fp32, head dim 32) ------------------------------------------------
// Two kernels, both tile the "other" side through shared memory (32 rows of 32 floats per operand):
//   stats + dq : one thread per query row of a (sample, head): pass 1 recomputes the softmax statistics (row maximum and
//                sum -> log-sum-exp), pass 2 accumulates dq = scale * sum_j ds_ij k_j with ds = p (dp - delta),
//                dp = <dout_i, v_j>, delta = <dout_i, o_i>; lse and delta are written to the workspace for
//   dk / dv    : one thread per key row: dv_j = sum_i p_ij dout_i, dk_j = scale * sum_i ds_ij q_i.
// nn.MultiheadAttention scales q before QK^T: s_ij = <scale * q_i, k_j>.
struct DenseBwdParams {
  const float* q; const float* k; const float* v; const float* o; const float* dout;
  long long ldq, ldk, ldv, ldo, ldd, qbs, kbs, vbs, obs, dbs;
  int B, Lq, Lk, heads;
  float scale;
  float* dq; float* dk; float* dv; float* lse; float* delta;
  DropoutRng rng;
};
constexpr int kBwdRows = 128, kBwdTile = 32, kHD = 32;

__global__ void __launch_bounds__(kBwdRows) attn_dense_bwd_dq_kernel(const DenseBwdParams p) {
  __shared__ float sk[kBwdTile][kHD + 1];
  __shared__ float sv[kBwdTile][kHD + 1];
  const int i = blockIdx.x * kBwdRows + threadIdx.x, h = blockIdx.y, b = blockIdx.z;
  const bool ok = i < p.Lq;
  float q[kHD], go[kHD], acc[kHD];
  float delta = 0.f;
#pragma unroll
  for (int d = 0; d < kHD; ++d) {
    q[d] = ok ? p.q[(long long)b * p.qbs + (long long)i * p.ldq + h * kHD + d] * p.scale : 0.f;
    go[d] = ok ? p.dout[(long long)b * p.dbs + (long long)i * p.ldd + h * kHD + d] : 0.f;
    const float ov = ok ? p.o[(long long)b * p.obs + (long long)i * p.ldo + h * kHD + d] : 0.f;
    delta = fmaf(go[d], ov, delta);
    acc[d] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int pass = 0; pass < 2; ++pass) {
    const float lse = m + logf(l);                       // valid in pass 1
    for (int j0 = 0; j0 < p.Lk; j0 += kBwdTile) {
      __syncthreads();
      for (int e = threadIdx.x; e < kBwdTile * kHD; e += kBwdRows) {
        const int r = e / kHD, d = e % kHD, j = j0 + r;
        sk[r][d] = j < p.Lk ? p.k[(long long)b * p.kbs + (long long)j * p.ldk + h * kHD + d] : 0.f;
        sv[r][d] = j < p.Lk ? p.v[(long long)b * p.vbs + (long long)j * p.ldv + h * kHD + d] : 0.f;
      }
      __syncthreads();
      const int nj = min(kBwdTile, p.Lk - j0);
      for (int r = 0; r < nj; ++r) {
        float sc = 0.f;
#pragma unroll
        for (int d = 0; d < kHD; ++d) sc = fmaf(q[d], sk[r][d], sc);
        if (pass == 0) {
          const float mn = fmaxf(m, sc);
          l = l * expf(m - mn) + expf(sc - mn);
          m = mn;
        } else {
          const float pij = expf(sc - lse);
          float dp = 0.f;
#pragma unroll
          for (int d = 0; d < kHD; ++d) dp = fmaf(go[d], sv[r][d], dp);
          if (p.rng.p > 0.f) dp *= dropout_scale(p.rng, (uint32_t)((b * p.heads + h) * p.Lq + i), (uint32_t)(j0 + r));
          const float ds = pij * (dp - delta);
#pragma unroll
          for (int d = 0; d < kHD; ++d) acc[d] = fmaf(ds, sk[r][d], acc[d]);
        }
      }
    }
  }
  if (!ok) return;
  const long long row = ((long long)b * p.heads + h) * p.Lq + i;
  p.lse[row] = m + logf(l);
  p.delta[row] = delta;
#pragma unroll
  for (int d = 0; d < kHD; ++d) p.dq[((long long)b * p.Lq + i) * (p.heads * kHD) + h * kHD + d] = acc[d] * p.scale;
}

__global__ void __launch_bounds__(kBwdRows) attn_dense_bwd_dkv_kernel(const DenseBwdParams p) {
  __shared__ float sq[kBwdTile][kHD + 1];
  __shared__ float sg[kBwdTile][kHD + 1];
  __shared__ float s_lse[kBwdTile], s_delta[kBwdTile];
  const int j = blockIdx.x * kBwdRows + threadIdx.x, h = blockIdx.y, b = blockIdx.z;
  const bool ok = j < p.Lk;
  float k[kHD], v[kHD], dk[kHD], dv[kHD];
#pragma unroll
  for (int d = 0; d < kHD; ++d) {
    k[d] = ok ? p.k[(long long)b * p.kbs + (long long)j * p.ldk + h * kHD + d] : 0.f;
    v[d] = ok ? p.v[(long long)b * p.vbs + (long long)j * p.ldv + h * kHD + d] : 0.f;
    dk[d] = 0.f; dv[d] = 0.f;
  }
  for (int i0 = 0; i0 < p.Lq; i0 += kBwdTile) {
    __syncthreads();
    for (int e = threadIdx.x; e < kBwdTile * kHD; e += kBwdRows) {
      const int r = e / kHD, d = e % kHD, i = i0 + r;
      sq[r][d] = i < p.Lq ? p.q[(long long)b * p.qbs + (long long)i * p.ldq + h * kHD + d] * p.scale : 0.f;
      sg[r][d] = i < p.Lq ? p.dout[(long long)b * p.dbs + (long long)i * p.ldd + h * kHD + d] : 0.f;
    }
    if (threadIdx.x < kBwdTile) {
      const int i = i0 + threadIdx.x;
      const long long row = ((long long)b * p.heads + h) * p.Lq + i;
      s_lse[threadIdx.x] = i < p.Lq ? p.lse[row] : INFINITY;          // exp(s - inf) = 0: padding rows contribute nothing
      s_delta[threadIdx.x] = i < p.Lq ? p.delta[row] : 0.f;
    }
    __syncthreads();
    for (int r = 0; r < kBwdTile; ++r) {
      float sc = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < kHD; ++d) { sc = fmaf(sq[r][d], k[d], sc); dp = fmaf(sg[r][d], v[d], dp); }
      const float pij = expf(sc - s_lse[r]);
      const float drop = (p.rng.p > 0.f && ok && i0 + r < p.Lq)
                             ? dropout_scale(p.rng, (uint32_t)((b * p.heads + h) * p.Lq + i0 + r), (uint32_t)j) : 1.0f;
      const float ds = pij * (dp * drop - s_delta[r]);
      const float pd = pij * drop;
#pragma unroll
      for (int d = 0; d < kHD; ++d) { dv[d] = fmaf(pd, sg[r][d], dv[d]); dk[d] = fmaf(ds, sq[r][d], dk[d]); }   // sq already holds scale * q
    }
  }
  if (!ok) return;
#pragma unroll
  for (int d = 0; d < kHD; ++d) {
    p.dk[((long long)b * p.Lk + j) * (p.heads * kHD) + h * kHD + d] = dk[d];
    p.dv[((long long)b * p.Lk + j) * (p.heads * kHD) + h * kHD + d] = dv[d];
  }
}

// TC_PW_* modes (include/transcar_b200.h)
__global__ void pointwise_kernel(const float* __restrict__ grad, const float* __restrict__ x, float* __restrict__ out, int n,
                                 int mode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float xv = x[i];
  if (mode == TC_PW_LOGIT) { out[i] = logit_f32(xv); return; }
  if (mode == TC_PW_SIGMOID) { out[i] = sigmoid_f32(xv); return; }
  const float g = grad[i];
  if (mode == TC_PW_SIGMOID_BWD) { out[i] = g * xv * (1.0f - xv); return; }
  const float eps = 1e-5f;
  float d = 0.f;
  if (xv >= 0.f && xv <= 1.f) {
    const float a = fmaxf(xv, eps), bq = fmaxf(1.0f - xv, eps);
    if (xv >= eps) d += 1.0f / a;
    if (1.0f - xv >= eps) d += 1.0f / bq;
  }
  out[i] = g * d;
}

__global__ void dropout_kernel(const float* __restrict__ x, const float* __restrict__ residual, float* __restrict__ out, int M,
                               int N4, const DropoutRng rng) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;        // one thread = 4 consecutive columns = one Philox block
  if (i >= (long long)M * N4) return;
  const int m = (int)(i / N4), c4 = (int)(i % N4);
  const uint4 r = philox4((uint32_t)c4, (uint32_t)m, rng);
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  float4 o = residual ? reinterpret_cast<const float4*>(residual)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  o.x += keep_from(r.x, rng.p) ? v.x * rng.inv_keep : 0.f;
  o.y += keep_from(r.y, rng.p) ? v.y * rng.inv_keep : 0.f;
  o.z += keep_from(r.z, rng.p) ? v.z * rng.inv_keep : 0.f;
  o.w += keep_from(r.w, rng.p) ? v.w * rng.inv_keep : 0.f;
  reinterpret_cast<float4*>(out)[i] = o;
}

__global__ void add_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n4,
                                int N4, int period) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const long long m = i / N4;
  const int c = (int)(i % N4);
  const float4 x = reinterpret_cast<const float4*>(a)[i];
  const float4 y = reinterpret_cast<const float4*>(b)[(m % period) * N4 + c];
  reinterpret_cast<float4*>(out)[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
}

__global__ void period_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int batches, long long pn) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pn) return;
  float s = 0.f;
  for (int b = 0; b < batches; ++b) s += x[(long long)b * pn + i];
  out[i] += s;
}

}  // namespace
}  // namespace tc

extern "C" int tc_attention_dense_bwd(const tc_attention_dense_bwd_args* a, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(a != nullptr, TC_ERR_NULL, "tc_attention_dense_bwd: args is NULL");
  TC_REQUIRE(a->q && a->k && a->v && a->o && a->dout && a->dq && a->dk && a->dv && a->workspace, TC_ERR_NULL,
             "tc_attention_dense_bwd: NULL pointer");
  TC_REQUIRE(a->D == 32, TC_ERR_SHAPE, "tc_attention_dense_bwd: head dim must be 32 (got %d)", a->D);
  TC_REQUIRE(a->B >= 0 && a->Lq >= 0 && a->Lk > 0 && a->heads > 0 && a->B <= 65535 && a->heads <= 65535, TC_ERR_SHAPE,
             "tc_attention_dense_bwd: bad shape");
  if (a->B == 0 || a->Lq == 0) return TC_OK;
  DenseBwdParams p{a->q, a->k, a->v, a->o, a->dout, a->ldq, a->ldk, a->ldv, a->ldo, a->ld_dout,
                   a->q_batch_stride, a->k_batch_stride, a->v_batch_stride, a->o_batch_stride, a->dout_batch_stride,
                   a->B, a->Lq, a->Lk, a->heads, a->scale, a->dq, a->dk, a->dv,
                   a->workspace, a->workspace + (long long)a->B * a->heads * a->Lq,
                   make_rng(a->dropout_p, a->dropout_seed, a->dropout_stream)};
  cudaStream_t s = as_stream(stream);
  attn_dense_bwd_dq_kernel<<<dim3((a->Lq + kBwdRows - 1) / kBwdRows, a->heads, a->B), kBwdRows, 0, s>>>(p);
  attn_dense_bwd_dkv_kernel<<<dim3((a->Lk + kBwdRows - 1) / kBwdRows, a->heads, a->B), kBwdRows, 0, s>>>(p);
  count_launch(2);
  return check_launch("tc_attention_dense_bwd");
}

extern "C" int tc_pointwise(const float* grad, const float* x, float* out, int32_t n, int32_t mode, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(x && out, TC_ERR_NULL, "tc_pointwise: NULL pointer");
  TC_REQUIRE(n >= 0 && mode >= TC_PW_LOGIT_BWD && mode <= TC_PW_SIGMOID, TC_ERR_SHAPE, "tc_pointwise: bad n / mode");
  TC_REQUIRE(grad || mode >= TC_PW_LOGIT, TC_ERR_NULL, "tc_pointwise: backward modes need grad");
  if (n == 0) return TC_OK;
  pointwise_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(grad, x, out, n, mode);
  count_launch();
  return check_launch("tc_pointwise");
}

extern "C" int tc_dropout(const float* x, const float* residual, float* out, int32_t M, int32_t N, float p, uint64_t seed,
                          uint64_t stream, tc_stream_t s) {
  using namespace tc;
  TC_REQUIRE(x && out, TC_ERR_NULL, "tc_dropout: NULL pointer");
  TC_REQUIRE(M >= 0 && N > 0 && N % 4 == 0, TC_ERR_SHAPE, "tc_dropout: N must be a positive multiple of 4 (got %d)", N);
  TC_REQUIRE(p >= 0.f && p < 1.f, TC_ERR_SHAPE, "tc_dropout: p must be in [0, 1)");
  TC_REQUIRE(aligned16(x) && aligned16(out) && (!residual || aligned16(residual)), TC_ERR_ALIGN, "tc_dropout: 16-byte alignment");
  if (M == 0) return TC_OK;
  const long long n4 = (long long)M * (N / 4);
  dropout_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, as_stream(s)>>>(x, residual, out, M, N / 4, make_rng(p, seed, stream));
  count_launch();
  return check_launch("tc_dropout");
}

extern "C" int tc_add_rows(const float* a, const float* b, float* out, int32_t M, int32_t N, int32_t period, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(a && b && out, TC_ERR_NULL, "tc_add_rows: NULL pointer");
  TC_REQUIRE(M >= 0 && N > 0 && N % 4 == 0 && period > 0, TC_ERR_SHAPE, "tc_add_rows: bad shape (N must be a multiple of 4)");
  TC_REQUIRE(aligned16(a) && aligned16(b) && aligned16(out), TC_ERR_ALIGN, "tc_add_rows: pointers must be 16-byte aligned");
  if (M == 0) return TC_OK;
  const long long n4 = (long long)M * (N / 4);
  add_rows_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, as_stream(stream)>>>(a, b, out, n4, N / 4, period);
  count_launch();
  return check_launch("tc_add_rows");
}

extern "C" int tc_period_sum(const float* x, float* out, int32_t batches, int32_t period, int32_t N, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(x && out, TC_ERR_NULL, "tc_period_sum: NULL pointer");
  TC_REQUIRE(batches >= 0 && period >= 0 && N > 0, TC_ERR_SHAPE, "tc_period_sum: bad shape");
  const long long pn = (long long)period * N;
  if (batches == 0 || pn == 0) return TC_OK;
  period_sum_kernel<<<(unsigned)((pn + 255) / 256), 256, 0, as_stream(stream)>>>(x, out, batches, pn);
  count_launch();
  return check_launch("tc_period_sum");
}

extern "C" int tc_transpose(const void* src, int32_t src_dtype, int64_t ld_src, void* dst, int32_t dst_dtype, int64_t ld_dst,
                            int32_t rows, int32_t cols, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(src && dst, TC_ERR_NULL, "tc_transpose: NULL pointer");
  TC_REQUIRE(rows >= 0 && cols >= 0 && ld_src >= cols && ld_dst >= rows, TC_ERR_SHAPE, "tc_transpose: bad shape");
  cudaStream_t s = as_stream(stream);
  if (dst_dtype == TC_BF16X2) {            // split output [cols, 2 * rows_pad], rows_pad = ld_dst / 2 (zero padded)
    TC_REQUIRE(src_dtype == TC_F32, TC_ERR_DTYPE, "tc_transpose: a split (TC_BF16X2) destination needs an fp32 source");
    TC_REQUIRE(ld_dst % 2 == 0 && ld_dst / 2 >= rows, TC_ERR_SHAPE, "tc_transpose: split destination needs ld_dst = 2 * rows_pad >= 2 * rows");
    if (cols == 0 || ld_dst == 0) return TC_OK;
    const int rows_pad = (int)(ld_dst / 2);
    launch(transpose_split_kernel, dim3((cols + 31) / 32, (rows_pad + 31) / 32), dim3(256), 0, s, 1u, (const float*)src,
           (long long)ld_src, (__nv_bfloat16*)dst, (long long)ld_dst, rows, rows_pad, cols);
    count_launch();
    return check_launch("tc_transpose(split)");
  }
  TC_REQUIRE((src_dtype == TC_F32 || src_dtype == TC_BF16) && (dst_dtype == TC_F32 || dst_dtype == TC_BF16), TC_ERR_DTYPE,
             "tc_transpose: bad dtype");
  if (rows == 0 || cols == 0) return TC_OK;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  const bool si = src_dtype == TC_BF16, di = dst_dtype == TC_BF16;
  if (!si && !di) launch(transpose_kernel<float, float>, grid, dim3(256), 0, s, 1u, (const float*)src, (long long)ld_src, (float*)dst, (long long)ld_dst, rows, cols);
  else if (!si && di) launch(transpose_kernel<float, __nv_bfloat16>, grid, dim3(256), 0, s, 1u, (const float*)src, (long long)ld_src, (__nv_bfloat16*)dst, (long long)ld_dst, rows, cols);
  else if (si && !di) launch(transpose_kernel<__nv_bfloat16, float>, grid, dim3(256), 0, s, 1u, (const __nv_bfloat16*)src, (long long)ld_src, (float*)dst, (long long)ld_dst, rows, cols);
  else launch(transpose_kernel<__nv_bfloat16, __nv_bfloat16>, grid, dim3(256), 0, s, 1u, (const __nv_bfloat16*)src, (long long)ld_src, (__nv_bfloat16*)dst, (long long)ld_dst, rows, cols);
  count_launch();
  return check_launch("tc_transpose");
}

extern "C" int tc_colsum(const void* x, int32_t dtype, int64_t ldx, int32_t M, int32_t N, float* out, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(x && out, TC_ERR_NULL, "tc_colsum: NULL pointer");
  TC_REQUIRE(M >= 0 && N > 0 && ldx >= N, TC_ERR_SHAPE, "tc_colsum: bad shape");
  TC_REQUIRE(dtype == TC_F32 || dtype == TC_BF16, TC_ERR_DTYPE, "tc_colsum: bad dtype");
  if (M == 0) return TC_OK;
  int slices = (M + 255) / 256;
  if (slices > 64) slices = 64;
  dim3 grid((N + 31) / 32, slices);
  if (dtype == TC_F32) launch(colsum_kernel<float>, grid, dim3(256), 0, as_stream(stream), 1u, (const float*)x, (long long)ldx, M, N, out);
  else launch(colsum_kernel<__nv_bfloat16>, grid, dim3(256), 0, as_stream(stream), 1u, (const __nv_bfloat16*)x, (long long)ldx, M, N, out);
  count_launch();
  return check_launch("tc_colsum");
}

extern "C" int tc_layernorm_fwd(const tc_layernorm_args* a, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(a != nullptr, TC_ERR_NULL, "tc_layernorm_fwd: args is NULL");
  TC_REQUIRE(a->x && a->gamma && a->beta && (a->y_f32 || a->y_bf16), TC_ERR_NULL, "tc_layernorm_fwd: NULL pointer");
  TC_REQUIRE(a->M >= 0 && a->N > 0 && a->N <= 32 * kLnMaxPerLane && a->ldx >= a->N && a->ldy >= a->N, TC_ERR_SHAPE,
             "tc_layernorm_fwd: bad shape");
  if (a->M == 0) return TC_OK;
  launch(layernorm_fwd_kernel, dim3((a->M + 7) / 8), dim3(256), 0, as_stream(stream), 1u, *a);
  count_launch();
  return check_launch("tc_layernorm_fwd");
}

extern "C" int tc_layernorm_bwd(const tc_layernorm_bwd_args* a, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(a != nullptr, TC_ERR_NULL, "tc_layernorm_bwd: args is NULL");
  TC_REQUIRE(a->dy && a->x && a->mean && a->rstd && a->gamma && a->dx, TC_ERR_NULL, "tc_layernorm_bwd: NULL pointer");
  TC_REQUIRE(a->M >= 0 && a->N > 0 && a->N <= 32 * kLnMaxPerLane, TC_ERR_SHAPE, "tc_layernorm_bwd: bad shape");
  if (a->M == 0) return TC_OK;
  int ctas = (a->M + 7) / 8;
  if (ctas > 296) ctas = 296;
  launch(layernorm_bwd_kernel, dim3(ctas), dim3(256), (size_t)2 * a->N * sizeof(float), as_stream(stream), 1u, *a);
  count_launch();
  return check_launch("tc_layernorm_bwd");
}

extern "C" int tc_mask_grad(const float* dy, int64_t ld_dy, const float* y, int64_t ld_y, const uint8_t* gate, float* dz,
                            int64_t ld_dz, int32_t M, int32_t N, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(dy && dz, TC_ERR_NULL, "tc_mask_grad: NULL pointer");
  TC_REQUIRE(M >= 0 && N > 0, TC_ERR_SHAPE, "tc_mask_grad: bad shape");
  if (M == 0) return TC_OK;
  const long long n = (long long)M * N;
  launch(mask_grad_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, as_stream(stream), 1u, dy, (long long)ld_dy, y,
         (long long)ld_y, gate, dz, (long long)ld_dz, M, N);
  count_launch();
  return check_launch("tc_mask_grad");
}

extern "C" int tc_attention_sparse_bwd(const tc_attention_bwd_args* a, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(a != nullptr, TC_ERR_NULL, "tc_attention_sparse_bwd: args is NULL");
  TC_REQUIRE(a->q && a->k && a->v && a->dout && a->geom && a->key_xy && a->dq && a->dk && a->dv, TC_ERR_NULL,
             "tc_attention_sparse_bwd: NULL pointer");
  TC_REQUIRE(a->heads == 8 && a->D == 32, TC_ERR_SHAPE, "tc_attention_sparse_bwd: needs 8 heads x 32");
  TC_REQUIRE(a->B >= 0 && a->Lq >= 0 && a->Lk >= 0 && a->B <= 65535, TC_ERR_SHAPE, "tc_attention_sparse_bwd: bad shape");
  TC_REQUIRE(aligned16(a->q) && aligned16(a->k) && aligned16(a->v) && aligned16(a->dout) && aligned16(a->dq) &&
                 a->ldq % 4 == 0 && a->ldk % 4 == 0 && a->ldv % 4 == 0 && a->ld_dout % 4 == 0 && a->ld_dq % 4 == 0 &&
                 a->q_batch_stride % 4 == 0 && a->k_batch_stride % 4 == 0 && a->v_batch_stride % 4 == 0,
             TC_ERR_ALIGN, "tc_attention_sparse_bwd: rows must be 16-byte aligned");
  if (a->B == 0 || a->Lq == 0) return TC_OK;
  SparseBwdParams p;
  p.q = a->q; p.k = a->k; p.v = a->v; p.dout = a->dout;
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv; p.ldo = a->ld_dout;
  p.qbs = a->q_batch_stride; p.kbs = a->k_batch_stride; p.vbs = a->v_batch_stride;
  p.B = a->B; p.Lq = a->Lq; p.Lk = a->Lk; p.scale = a->scale;
  p.geom = a->geom; p.key_xy = a->key_xy;
  p.dq = a->dq; p.ld_dq = a->ld_dq;
  p.dk = a->dk; p.dv = a->dv; p.ld_dk = a->ld_dk; p.ld_dv = a->ld_dv;
  p.dkbs = a->dk_batch_stride; p.dvbs = a->dv_batch_stride;
  p.rng = make_rng(a->dropout_p, a->dropout_seed, a->dropout_stream);
  dim3 grid((a->Lq + kWarps - 1) / kWarps, a->B);
  launch(attention_sparse_bwd_kernel, grid, dim3(kWarps * 32), 0, as_stream(stream), 1u, p);
  count_launch();
  return check_launch("tc_attention_sparse_bwd");
}
