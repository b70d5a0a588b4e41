// K4 (radar path): softmax(scale * Q K^T + distance mask) V exploiting the sparsity of the TransCAR mask.
//
// The radar mask of detr3d_head.py:549-571 lets a query attend only to radar points within 0.5-2 m of three
// circle centres; with ~1.5k points spread over a 102 m x 102 m range that is ~0.1 % of the [Q, R] pairs
// (~260 of 900 queries have any key at all, most of those one to five).  A dense-tile kernel
// (attention_tc.cu / attention_simt.cu) spends all of its time on pairs whose probability is exactly zero and
// re-derives the same mask once per head.  This kernel does the necessary work only:
//   * one warp per (sample, query), 8 queries per CTA, four CTAs per SM (64 registers: at 98 registers and 16 warps a
//     single CTA fitted an SM and the 456-CTA grid ran as three waves - 21.2 -> 16.1 us per launch); every lane owns 8
//     consecutive channels of the 256-wide embedding, so a head (32 channels) is a group of 4 lanes and all 8 heads share
//     ONE pass over the keys;
//   * scan: 32 keys per iteration, one key per lane, a conservative squared-distance prefilter (5 instructions);
//     only candidates run the exact test, which is bit-identical to torch.cdist's mm route (tc_common.cuh);
//   * for each allowed key (rare): one 512-byte K row and V row read (16 B per lane), per-head dot product reduced
//     over the 4-lane group, fp32 online softmax, fp32 accumulation.
// Rows without any allowed key produce zeros and row_any = 0 (quirk Q6).  Works for fp32 and bf16 operands.
#include "tc_common.cuh"

namespace tc {
namespace {

constexpr int kWarps = 8;

struct SparseParams {
  const void* q; const void* k; const void* v;
  long long ldq, ldk, ldv, qbs, kbs, vbs;
  int B, Lq, Lk;
  float scale;
  const float* geom; const float* key_xy;
  void* out; long long ldo;
  int out_split;          // 16-bit output is split bf16 [.., 2*256]: hi | lo
  uint8_t* row_any;
  DropoutRng rng;         // p == 0: no dropout
};

template <bool kBf16>
__device__ __forceinline__ void load8(const void* base, long long off, float (&d)[8]) {
  if (kBf16) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(base) + off));
    d[0] = bf16_lo(u.x); d[1] = bf16_hi(u.x); d[2] = bf16_lo(u.y); d[3] = bf16_hi(u.y);
    d[4] = bf16_lo(u.z); d[5] = bf16_hi(u.z); d[6] = bf16_lo(u.w); d[7] = bf16_hi(u.w);
  } else {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + off);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
  }
}

constexpr int kKeyTile = 2048;           // keys staged in shared memory per pass (x, y, |k|^2: 24 KB)

template <bool kBf16In, bool kBf16Out>
__global__ void __launch_bounds__(kWarps * 32, 4) attention_sparse_kernel(const SparseParams p) {
  __shared__ float2 s_kxy[kKeyTile];
  __shared__ float s_kn[kKeyTile];
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  const bool active = q < p.Lq;
  const long long row = (long long)b * p.Lq + (active ? q : 0);
  pdl_trigger();
  pdl_wait();

  const float* g = p.geom + row * 8;
  const float cx = __ldg(g + 0), cy = __ldg(g + 1), fx = __ldg(g + 2), fy = __ldg(g + 3);
  const float rx = __ldg(g + 4), ry = __ldg(g + 5), radius = __ldg(g + 6);
  const Circle cc = make_circle(cx, cy), cf = make_circle(fx, fy), cr = make_circle(rx, ry);
  // Prefilter radius: every allowed key lies within `radius` of c, f or r, hence within radius + max(|f-c|, |r-c|)
  // of c.  0.25 m of slack dwarfs the rounding of the mm-route distance (~1e-3 m at 72 m from the origin).
  // NaN/inf geometry makes the bound NaN/inf, and `!(d2 >= bound2)` then passes every key on to the exact test.
  const float span = fmaxf(sqrtf((fx - cx) * (fx - cx) + (fy - cy) * (fy - cy)),
                           sqrtf((rx - cx) * (rx - cx) + (ry - cy) * (ry - cy)));
  const float bound = radius + span + 0.25f;
  const float bound2 = bound * bound;

  float qv[8];
  load8<kBf16In>(p.q, (long long)b * p.qbs + (long long)(active ? q : 0) * p.ldq + lane * 8, qv);
#pragma unroll
  for (int i = 0; i < 8; ++i) qv[i] *= p.scale;            // nn.MultiheadAttention scales q before QK^T

  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;                    // per head (identical in the 4 lanes of a head group)

  const float2* keys = reinterpret_cast<const float2*>(p.key_xy) + (long long)b * p.Lk;
  for (int t0 = 0; t0 < p.Lk; t0 += kKeyTile) {
    const int nt = min(kKeyTile, p.Lk - t0);
    __syncthreads();
    const int ntp = (nt + 127) & ~127;
    for (int i = threadIdx.x; i < ntp; i += kWarps * 32) {  // the CTA's queries share one staged copy of the keys
      // the tail of the last 128-key group is padded with points whose squared distance overflows to +inf, so the scan
      // needs no bounds test
      const float2 xy = i < nt ? __ldg(keys + t0 + i) : make_float2(3e19f, 3e19f);
      s_kxy[i] = xy; s_kn[i] = key_norm(xy.x, xy.y);
    }
    __syncthreads();
    if (!active) continue;
    // 128 keys per iteration: four independent prefilters, ONE vote and branch.  (32 keys per iteration cost ~210
    // cycles each - shared-memory load -> 4 dependent fp32 ops -> vote -> branch, nothing to overlap - 10 000 cycles
    // for the 1 500 keys of a sample, measured with clock64.)
    for (int k128 = 0; k128 < ntp; k128 += 128) {
      bool cand4[4];
      bool any4 = false;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 kxy = s_kxy[k128 + u * 32 + lane];
        const float dx = kxy.x - cx, dy = kxy.y - cy;
        cand4[u] = !(fmaf(dx, dx, dy * dy) >= bound2);
        any4 |= cand4[u];
      }
      if (!__any_sync(0xffffffffu, any4)) continue;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
      const int k0 = k128 + u * 32;
      const int key = k0 + lane;
      if (!__any_sync(0xffffffffu, cand4[u])) continue;
      const float kx = s_kxy[key].x, ky = s_kxy[key].y;
      bool ok = false;
      if (cand4[u] && key < nt) ok = radar_allowed(cc, cf, cr, radius, kx, ky, s_kn[key]);
      unsigned todo = __ballot_sync(0xffffffffu, ok);
      while (todo) {
        const int j = t0 + k0 + __ffs(todo) - 1;
        todo &= todo - 1;
        float kv[8], vv[8];
        load8<kBf16In>(p.k, (long long)b * p.kbs + (long long)j * p.ldk + lane * 8, kv);
        load8<kBf16In>(p.v, (long long)b * p.vbs + (long long)j * p.ldv + lane * 8, vv);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(qv[i], kv[i], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);             // head score, same value in the 4 lanes of the head
        const float m_new = fmaxf(m_run, s);
        const float corr = expf(m_run - m_new);               // m_run = -inf -> 0
        const float pj = expf(s - m_new);
        l_run = fmaf(l_run, corr, pj);
        // attention-probability dropout (training): the normaliser keeps every key, the sum over V does not
        const float pd = p.rng.p > 0.f ? pj * dropout_scale(p.rng, (uint32_t)((b * 8 + (lane >> 2)) * p.Lq + q), (uint32_t)j) : pj;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(acc[i], corr, pd * vv[i]);
        m_run = m_new;
      }
      }
    }
  }
  if (!active) return;

  const bool any = l_run > 0.f;
  const float inv = any ? 1.0f / l_run : 0.f;
  const long long oo = row * p.ldo + lane * 8;
  if (kBf16Out) {
    uint4 u;
    u.x = pack_bf16(acc[0] * inv, acc[1] * inv); u.y = pack_bf16(acc[2] * inv, acc[3] * inv);
    u.z = pack_bf16(acc[4] * inv, acc[5] * inv); u.w = pack_bf16(acc[6] * inv, acc[7] * inv);
    *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + oo) = u;
    if (p.out_split) {
      uint4 l;
      l.x = pack_bf16(acc[0] * inv - bf16_lo(u.x), acc[1] * inv - bf16_hi(u.x));
      l.y = pack_bf16(acc[2] * inv - bf16_lo(u.y), acc[3] * inv - bf16_hi(u.y));
      l.z = pack_bf16(acc[4] * inv - bf16_lo(u.z), acc[5] * inv - bf16_hi(u.z));
      l.w = pack_bf16(acc[6] * inv - bf16_lo(u.w), acc[7] * inv - bf16_hi(u.w));
      *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + oo + 256) = l;
    }
  } else {
    float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.out) + oo);
    o[0] = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
    o[1] = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
  }
  if (p.row_any && lane == 0) p.row_any[row] = any ? 1 : 0;
}

}  // namespace

bool attention_sparse_supported(const tc_attention_args* a) {
  return a->geom != nullptr && a->heads * a->D == 256 && a->D == 32 && (a->qkv_dtype == TC_F32 || a->qkv_dtype == TC_BF16);
}

int attention_sparse_launch(const tc_attention_args* a, cudaStream_t s) {
  SparseParams p;
  p.q = a->q; p.k = a->k; p.v = a->v;
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv;
  p.qbs = a->q_batch_stride; p.kbs = a->k_batch_stride; p.vbs = a->v_batch_stride;
  p.B = a->B; p.Lq = a->Lq; p.Lk = a->Lk;
  p.scale = a->scale; p.geom = a->geom; p.key_xy = a->key_xy;
  p.out = a->out; p.ldo = a->ldo; p.row_any = a->row_any;
  p.out_split = a->out_dtype == TC_BF16X2 ? 1 : 0;
  p.rng = make_rng(a->dropout_p, a->dropout_seed, a->dropout_stream);
  dim3 grid((a->Lq + kWarps - 1) / kWarps, a->B);
  const bool bi = a->qkv_dtype == TC_BF16, bo = a->out_dtype == TC_BF16 || a->out_dtype == TC_BF16X2;
  if (bi && bo) launch(attention_sparse_kernel<true, true>, grid, dim3(kWarps * 32), 0, s, 1u, p);
  else if (bi) launch(attention_sparse_kernel<true, false>, grid, dim3(kWarps * 32), 0, s, 1u, p);
  else if (bo) launch(attention_sparse_kernel<false, true>, grid, dim3(kWarps * 32), 0, s, 1u, p);
  else launch(attention_sparse_kernel<false, false>, grid, dim3(kWarps * 32), 0, s, 1u, p);
  count_launch();
  return check_launch("tc_attention_fwd(sparse)");
}

}  // namespace tc
