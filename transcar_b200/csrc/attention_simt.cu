// K4 (exact path): softmax(scale * Q K^T + mask) V on the CUDA cores, fp32, online softmax.
// One thread per query row (q and the 32-wide accumulator stay in registers), K/V streamed through
// shared memory in 64-key tiles (every thread reads the same key -> shared-memory broadcast).
// The radar distance mask (detr3d_head.py:549-571) is evaluated in-kernel from per-query circle
// geometry; 8-key chunks with no allowed key are skipped, which is ~99.9 % of them on the radar path.
// The tensor-core (tcgen05) kernel is attention_tc.cu; this one is the fp32 parity mode.
#include "tc_common.cuh"

namespace tc {
namespace {

constexpr int kD = 32;
constexpr int kTile = 64;
constexpr int kRows = 128;

struct AttnParams {
  const void* q; const void* k; const void* v;
  long long ldq, ldk, ldv, qbs, kbs, vbs;
  int B, Lq, Lk, heads;
  float scale;
  const float* geom; const float* key_xy;
  void* out; long long ldo;
  uint8_t* row_any;
  DropoutRng rng;          // p == 0: no dropout
  const uint8_t* attn_blocked; const uint8_t* key_blocked;      // generic nn.MultiheadAttention masks (may be NULL)
};

template <bool kBf16>
__device__ __forceinline__ void load_row32(const void* base, long long off, float (&dst)[kD]) {
  if (kBf16) {
    const uint4* p = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(base) + off);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u = p[i];
      dst[i * 8 + 0] = bf16_lo(u.x); dst[i * 8 + 1] = bf16_hi(u.x);
      dst[i * 8 + 2] = bf16_lo(u.y); dst[i * 8 + 3] = bf16_hi(u.y);
      dst[i * 8 + 4] = bf16_lo(u.z); dst[i * 8 + 5] = bf16_hi(u.z);
      dst[i * 8 + 6] = bf16_lo(u.w); dst[i * 8 + 7] = bf16_hi(u.w);
    }
  } else {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + off);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 u = p[i];
      dst[i * 4 + 0] = u.x; dst[i * 4 + 1] = u.y; dst[i * 4 + 2] = u.z; dst[i * 4 + 3] = u.w;
    }
  }
}

template <bool kBf16In, bool kBf16Out, bool kMask>
__global__ void __launch_bounds__(kRows) attention_simt_kernel(const AttnParams p) {
  __shared__ __align__(16) float Ks[kTile][kD];
  __shared__ __align__(16) float Vs[kTile][kD];
  __shared__ float Kx[kTile], Ky[kTile], Kn[kTile];
  const int b = blockIdx.z, h = blockIdx.y;
  const int row = blockIdx.x * kRows + threadIdx.x;
  const bool active = row < p.Lq;

  float q[kD], acc[kD];
#pragma unroll
  for (int d = 0; d < kD; ++d) { q[d] = 0.f; acc[d] = 0.f; }
  if (active) {
    load_row32<kBf16In>(p.q, (long long)b * p.qbs + (long long)row * p.ldq + h * kD, q);
#pragma unroll
    for (int d = 0; d < kD; ++d) q[d] *= p.scale;       // reference scales q before QK^T
  }
  Circle cc, cf, cr;
  float radius = 0.f;
  if (kMask) {
    cc = make_circle(0.f, 0.f); cf = cc; cr = cc;
    if (active) {
      const float* g = p.geom + ((long long)b * p.Lq + row) * 8;
      cc = make_circle(g[0], g[1]); cf = make_circle(g[2], g[3]); cr = make_circle(g[4], g[5]);
      radius = g[6];
    }
  }
  float mrun = -INFINITY, lrun = 0.f;

  for (int k0 = 0; k0 < p.Lk; k0 += kTile) {
    __syncthreads();
    // cooperative tile load: 64 keys x 32 dims for K and V (8 threads per key, 4 dims each... 128 threads -> 2 passes x 2)
    for (int i = threadIdx.x; i < kTile * (kD / 4); i += kRows) {
      const int j = i / (kD / 4), d4 = (i % (kD / 4)) * 4;
      const int key = k0 + j;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (key < p.Lk) {
        const long long ko = (long long)b * p.kbs + (long long)key * p.ldk + h * kD + d4;
        const long long vo = (long long)b * p.vbs + (long long)key * p.ldv + h * kD + d4;
        if (kBf16In) {
          uint2 a = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(p.k) + ko);
          uint2 c = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(p.v) + vo);
          kv = make_float4(bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y));
          vv = make_float4(bf16_lo(c.x), bf16_hi(c.x), bf16_lo(c.y), bf16_hi(c.y));
        } else {
          kv = *reinterpret_cast<const float4*>(static_cast<const float*>(p.k) + ko);
          vv = *reinterpret_cast<const float4*>(static_cast<const float*>(p.v) + vo);
        }
      }
      *reinterpret_cast<float4*>(&Ks[j][d4]) = kv;
      *reinterpret_cast<float4*>(&Vs[j][d4]) = vv;
    }
    if (kMask) {
      for (int j = threadIdx.x; j < kTile; j += kRows) {
        const int key = k0 + j;
        float kx = 0.f, ky = 0.f;
        if (key < p.Lk) {
          kx = p.key_xy[((long long)b * p.Lk + key) * 2 + 0];
          ky = p.key_xy[((long long)b * p.Lk + key) * 2 + 1];
        }
        Kx[j] = kx; Ky[j] = ky; Kn[j] = key_norm(kx, ky);
      }
    }
    __syncthreads();
    if (!active) continue;
    const int nk = min(kTile, p.Lk - k0);
    for (int j0 = 0; j0 < nk; j0 += 8) {
      bool ok[8];
      bool any = false;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int jj = j0 + j;
        bool a = jj < nk;
        if (kMask && a) a = radar_allowed(cc, cf, cr, radius, Kx[jj], Ky[jj], Kn[jj]);
        if (a && p.attn_blocked) a = p.attn_blocked[(long long)row * p.Lk + k0 + jj] == 0;
        if (a && p.key_blocked) a = p.key_blocked[(long long)b * p.Lk + k0 + jj] == 0;
        ok[j] = a;
        any |= a;
      }
      if (!any) continue;
      float s[8];
      float cmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float dot = 0.f;
        if (ok[j]) {
          const float4* kr = reinterpret_cast<const float4*>(&Ks[j0 + j][0]);
#pragma unroll
          for (int d4 = 0; d4 < kD / 4; ++d4) {
            const float4 kk = kr[d4];
            dot = fmaf(q[d4 * 4 + 0], kk.x, dot); dot = fmaf(q[d4 * 4 + 1], kk.y, dot);
            dot = fmaf(q[d4 * 4 + 2], kk.z, dot); dot = fmaf(q[d4 * 4 + 3], kk.w, dot);
          }
        }
        s[j] = ok[j] ? dot : -INFINITY;
        cmax = fmaxf(cmax, s[j]);
      }
      const float mnew = fmaxf(mrun, cmax);
      const float corr = expf(mrun - mnew);             // mrun = -inf -> 0
      lrun *= corr;
#pragma unroll
      for (int d = 0; d < kD; ++d) acc[d] *= corr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (!ok[j]) continue;
        float pj = expf(s[j] - mnew);
        lrun += pj;
        if (p.rng.p > 0.f)           // attention-probability dropout: the normaliser keeps every key, the sum over V does not
          pj *= dropout_scale(p.rng, (uint32_t)((b * p.heads + h) * p.Lq + row), (uint32_t)(k0 + j0 + j));
        const float4* vr = reinterpret_cast<const float4*>(&Vs[j0 + j][0]);
#pragma unroll
        for (int d4 = 0; d4 < kD / 4; ++d4) {
          const float4 vv = vr[d4];
          acc[d4 * 4 + 0] = fmaf(pj, vv.x, acc[d4 * 4 + 0]); acc[d4 * 4 + 1] = fmaf(pj, vv.y, acc[d4 * 4 + 1]);
          acc[d4 * 4 + 2] = fmaf(pj, vv.z, acc[d4 * 4 + 2]); acc[d4 * 4 + 3] = fmaf(pj, vv.w, acc[d4 * 4 + 3]);
        }
      }
      mrun = mnew;
    }
  }
  if (!active) return;
  const float inv = lrun > 0.f ? 1.0f / lrun : 0.f;
  const long long oo = ((long long)b * p.Lq + row) * p.ldo + h * kD;
  if (kBf16Out) {
    uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + oo);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16(acc[i * 8 + 0] * inv, acc[i * 8 + 1] * inv); u.y = pack_bf16(acc[i * 8 + 2] * inv, acc[i * 8 + 3] * inv);
      u.z = pack_bf16(acc[i * 8 + 4] * inv, acc[i * 8 + 5] * inv); u.w = pack_bf16(acc[i * 8 + 6] * inv, acc[i * 8 + 7] * inv);
      o[i] = u;
    }
  } else {
    float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.out) + oo);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      o[i] = make_float4(acc[i * 4 + 0] * inv, acc[i * 4 + 1] * inv, acc[i * 4 + 2] * inv, acc[i * 4 + 3] * inv);
  }
  if (p.row_any && h == 0) p.row_any[(long long)b * p.Lq + row] = lrun > 0.f ? 1 : 0;
}

// ---- radar geometry + materialised mask -------------------------------------------------------------
__global__ void radar_geometry_kernel(const tc_radar_geometry_args a) {
  pdl_trigger();
  pdl_wait();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= a.M) return;
  float cx = a.centre[(long long)m * a.ld_centre + 0], cy = a.centre[(long long)m * a.ld_centre + 1];
  if (a.centre_is_normalised) {            // H:545-546
    cx = __fadd_rn(__fmul_rn(cx, a.pc_range[3] - a.pc_range[0]), a.pc_range[0]);
    cy = __fadd_rn(__fmul_rn(cy, a.pc_range[4] - a.pc_range[1]), a.pc_range[1]);
  }
  const float* code = a.code + (long long)m * a.ld_code;
  float g[8];
  radar_geom_row(cx, cy, code[3], code[6], code[7], a.r_lo, a.r_hi, g);
  float* o = a.geom + (long long)m * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = g[i];
}

__global__ void __launch_bounds__(256) radar_mask_kernel(const float* __restrict__ geom, const float* __restrict__ key_xy,
                                                         int Lq, int Lk, uint8_t* blocked, uint8_t* row_any) {
  // one warp per query row
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.y * Lq + blockIdx.x * 8 + (threadIdx.x >> 5);
  if (blockIdx.x * 8 + (threadIdx.x >> 5) >= Lq) return;
  const float* g = geom + row * 8;
  const Circle cc = make_circle(g[0], g[1]), cf = make_circle(g[2], g[3]), cr = make_circle(g[4], g[5]);
  const float radius = g[6];
  const float* kb = key_xy + (long long)blockIdx.y * Lk * 2;
  bool any = false;
  for (int k = lane; k < Lk; k += 32) {
    const float kx = kb[k * 2], ky = kb[k * 2 + 1];
    const bool ok = radar_allowed(cc, cf, cr, radius, kx, ky, key_norm(kx, ky));
    any |= ok;
    if (blocked) blocked[row * Lk + k] = ok ? 0 : 1;
  }
  any = __any_sync(0xffffffffu, any);
  if (row_any && lane == 0) row_any[row] = any ? 1 : 0;
}

}  // namespace

int attention_simt_launch(const tc_attention_args* a, cudaStream_t s) {
  AttnParams p;
  p.q = a->q; p.k = a->k; p.v = a->v;
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv;
  p.qbs = a->q_batch_stride; p.kbs = a->k_batch_stride; p.vbs = a->v_batch_stride;
  p.B = a->B; p.Lq = a->Lq; p.Lk = a->Lk; p.heads = a->heads;
  p.scale = a->scale; p.geom = a->geom; p.key_xy = a->key_xy;
  p.out = a->out; p.ldo = a->ldo; p.row_any = a->row_any;
  p.rng = make_rng(a->dropout_p, a->dropout_seed, a->dropout_stream);
  p.attn_blocked = a->attn_blocked; p.key_blocked = a->key_blocked;
  dim3 grid((a->Lq + kRows - 1) / kRows, a->heads, a->B);
  const bool bi = a->qkv_dtype == TC_BF16, bo = a->out_dtype == TC_BF16, mk = a->geom != nullptr;
#define TC_ATTN_LAUNCH(BI, BO, MK) attention_simt_kernel<BI, BO, MK><<<grid, kRows, 0, s>>>(p)
  if (bi) { if (bo) { if (mk) TC_ATTN_LAUNCH(true, true, true); else TC_ATTN_LAUNCH(true, true, false); }
            else    { if (mk) TC_ATTN_LAUNCH(true, false, true); else TC_ATTN_LAUNCH(true, false, false); } }
  else    { if (bo) { if (mk) TC_ATTN_LAUNCH(false, true, true); else TC_ATTN_LAUNCH(false, true, false); }
            else    { if (mk) TC_ATTN_LAUNCH(false, false, true); else TC_ATTN_LAUNCH(false, false, false); } }
#undef TC_ATTN_LAUNCH
  count_launch();
  return check_launch("tc_attention_fwd(simt)");
}

}  // namespace tc

extern "C" int tc_radar_geometry(const tc_radar_geometry_args* a, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(a != nullptr, TC_ERR_NULL, "tc_radar_geometry: args is NULL");
  TC_REQUIRE(a->centre && a->code && a->geom, TC_ERR_NULL, "tc_radar_geometry: NULL pointer");
  TC_REQUIRE(a->M >= 0 && a->ld_centre >= 2 && a->ld_code >= 8, TC_ERR_SHAPE, "tc_radar_geometry: bad shape");
  if (a->M == 0) return TC_OK;
  launch(radar_geometry_kernel, dim3((a->M + 127) / 128), dim3(128), 0, as_stream(stream), 1u, *a);
  count_launch();
  return check_launch("tc_radar_geometry");
}

extern "C" int tc_radar_mask(const float* geom, const float* key_xy, int32_t B, int32_t Lq, int32_t Lk,
                             uint8_t* blocked, uint8_t* row_any, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(geom && key_xy, TC_ERR_NULL, "tc_radar_mask: NULL pointer");
  TC_REQUIRE(blocked || row_any, TC_ERR_NULL, "tc_radar_mask: no output");
  TC_REQUIRE(B >= 0 && Lq >= 0 && Lk >= 0 && B <= 65535, TC_ERR_SHAPE, "tc_radar_mask: bad shape");
  if (B == 0 || Lq == 0) return TC_OK;
  dim3 grid((Lq + 7) / 8, B);
  radar_mask_kernel<<<grid, 256, 0, as_stream(stream)>>>(geom, key_xy, Lq, Lk, blocked, row_any);
  count_launch();
  return check_launch("tc_radar_mask");
}
