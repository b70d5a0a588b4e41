// K3m fused three-layer head on tcgen05 (bf16x3 operand format):
//     Y = W3 f2(W2 f1(W1 X + b1) + b2) + b3,   f = ReLU or ReLU(LayerNorm(.)),   C = 256 wide, N3 <= 32 outputs,
// followed by the row-local tails of tc_linear (reference-point update / box anchor + next radar-mask geometry).
//
// Replaces the three dependent tc_linear launches of the refinement branches (reg_branches.l.{0,2,4}: built at H:208-213, used
// at T:190-203), of the radar head's regression heads (final_reg*.{0,2,4}: H:84-90 / 102-108 / 120-126, used at H:593-600,
// H:662-665, H:720-723) and classification heads (final_cls*.{0,1,3,4,6}: Linear + LayerNorm + ReLU twice, then Linear;
// H:74-83 / 92-101 / 110-119, used at H:592, H:661, H:719).  A 128-row block stays inside one CTA: the hidden
// activations live in tensor memory (two 256-column accumulators), are turned - bias, optional LayerNorm over the row, ReLU,
// hi / lo split - 64 columns at a time into the swizzled K-major A tiles of the next GEMM in shared memory, and only the
// [M, N3] result leaves the SM.  Unfused, each hidden activation made a round trip through L2 in split form (7.4 MB written,
// re-read four times) and every launch paid its dependency wait, first-tile latency and store tail: ~7 + 7 + 6 us back to
// back, 45-55 us inside the step where the chains of a radar layer run beside each other.
// Ring: 2 stages x (A_hi | A_lo | W_hi | W_lo) = 2 x 96 KB; iterations 0..3 = GEMM 1 (X and W1 tiles by TMA), 4..7 = GEMM 2,
// 8..11 = GEMM 3 (A tile written by the conversion warps, W tile by TMA).
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = conversion + epilogue.
#include <cuda.h>

#include <cstdlib>

#include "tc_common.cuh"
#include "tc_epilogue.cuh"
#include "tc_sm100.cuh"

namespace tc {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 320;
constexpr int C = 256;                  // width of the input and of both hidden layers
constexpr int N3P = 32;                 // padded width of the last layer (one tcgen05.ld chunk)
constexpr int kKB = C / BK;
constexpr uint32_t kATile = BM * BK * 2;                       // 16 KB
constexpr uint32_t kWTile = 256 * BK * 2;                      // 32 KB
constexpr uint32_t kW3Tile = N3P * BK * 2;                     // 4 KB
constexpr uint32_t kStageBytes = 2 * kATile + 2 * kWTile;      // 96 KB
constexpr int kStages = 2;
constexpr uint32_t kOffBars = kStages * kStageBytes;           // full[2], a2[2], empty[2], acc[3], tmem slot
constexpr uint32_t kOffVec = kOffBars + 128;                   // b1, g1, be1, b2, g2, be2 (C each), b3 (N3P)
constexpr uint32_t kOffStat = kOffVec + (6 * C + N3P) * 4;     // LayerNorm partials: 2 layers x 2 halves x BM x (sum, sumsq)
constexpr uint32_t kSmemUsed = kOffStat + 2 * 2 * BM * 8;
constexpr size_t kSmemBytes = kSmemUsed + 1024;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

struct MlpParams {
  int M, N3;
  const float* b1; const float* g1; const float* be1;     // g* == null: no LayerNorm in that layer
  const float* b2; const float* g2; const float* be2;
  const float* b3;
  float ln_eps;
  float* out; long long ld_out;
  TailParams tail;
  int w_static;
  unsigned long long* trace;
};

__global__ void __launch_bounds__(kThreads, 1)
mlp_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
              const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_w3, const MlpParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* s_vec = reinterpret_cast<float*>(smem + kOffVec);
  float* s_b3 = s_vec + 6 * C;
  float2* s_stat = reinterpret_cast<float2*>(smem + kOffStat);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), a20 = smem_u32(bars + 2), empty0 = smem_u32(bars + 4), acc0 = smem_u32(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
#ifdef TC_TRACE_BUILD
  unsigned long long* trc = p.trace ? p.trace + 16ull * blockIdx.x : nullptr;
  if (trc && threadIdx.x == 0) { trc[0] = gtime(); trc[10] = smid(); TC_TRACE(1); }
#endif
  pdl_trigger();

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(a20 + 8 * s, 8); mbar_init(empty0 + 8 * s, 1); }
    for (int l = 0; l < 3; ++l) mbar_init(acc0 + 8 * l, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w3) : "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);        // two 256-column accumulators (the third reuses columns 0..31)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // static weights: the W1 tiles of the first two stages are requested before the dependency wait (see linear_tc.cu)
  const int num_pre = p.w_static ? kStages : 0;
  if (threadIdx.x == 0) {
    for (int it = 0; it < num_pre; ++it) {
      const uint32_t w_dst = smem_base + it * kStageBytes + 2 * kATile;
      mbar_expect_tx(full0 + 8 * it, kStageBytes);
      tma_load_2d(w_dst, &map_w1, it * BK, 0, full0 + 8 * it);
      tma_load_2d(w_dst + kWTile, &map_w1, C + it * BK, 0, full0 + 8 * it);
    }
  }
  pdl_wait();            // everything above touched parameters only; activations are read from here on
  if (threadIdx.x == 0) TC_TRACE(3);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int it = 0; it < 3 * kKB; ++it) {
        const int s = it & 1, layer = it / kKB, kb = it % kKB;
        const uint32_t ph = (it >> 1) & 1;
        const uint32_t st = smem_base + s * kStageBytes, full = full0 + 8 * s;
        const bool w_done = it < num_pre;
        if (!w_done) {
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          mbar_expect_tx(full, layer == 0 ? kStageBytes : layer == 1 ? 2 * kWTile : 2 * kW3Tile);
        }
        if (layer == 0) {
          tma_load_2d(st, &map_x, kb * BK, m0, full);
          tma_load_2d(st + kATile, &map_x, C + kb * BK, m0, full);
          if (!w_done) {
            tma_load_2d(st + 2 * kATile, &map_w1, kb * BK, 0, full);
            tma_load_2d(st + 2 * kATile + kWTile, &map_w1, C + kb * BK, 0, full);
          }
        } else if (layer == 1) {
          tma_load_2d(st + 2 * kATile, &map_w2, kb * BK, 0, full);
          tma_load_2d(st + 2 * kATile + kWTile, &map_w2, C + kb * BK, 0, full);
        } else {
          tma_load_2d(st + 2 * kATile, &map_w3, kb * BK, 0, full);
          tma_load_2d(st + 2 * kATile + kWTile, &map_w3, C + kb * BK, 0, full);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer: UMMA 128 x 256 x 16 (layers 1, 2) / 128 x 32 x 16 (layer 3), three passes per K step =====
    if (lane == 0) {
      constexpr uint32_t idesc_w = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      constexpr uint32_t idesc_3 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N3P >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int it = 0; it < 3 * kKB; ++it) {
        const int s = it & 1, layer = it / kKB, kb = it % kKB;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(full0 + 8 * s, ph);
        if (layer > 0) mbar_wait(a20 + 8 * s, (uint32_t)((it - kKB) >> 1) & 1);
        tc_fence_after();
        if (it == 0) TC_TRACE(4);
        if (it == kKB) TC_TRACE(5);
        if (it == 2 * kKB) TC_TRACE(6);
        const uint32_t st = smem_base + s * kStageBytes;
        const uint32_t d = tmem_base + (layer == 1 ? 256u : 0u);
        const uint32_t idesc = layer == 2 ? idesc_3 : idesc_w;
        const uint64_t dah = make_desc_sw128(st), dal = make_desc_sw128(st + kATile);
        const uint64_t dwh = make_desc_sw128(st + 2 * kATile), dwl = make_desc_sw128(st + 2 * kATile + kWTile);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(d, dah + 2 * k, dwh + 2 * k, idesc, (kb == 0 && k == 0) ? 0u : 1u);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(d, dal + 2 * k, dwh + 2 * k, idesc, 1u);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(d, dah + 2 * k, dwl + 2 * k, idesc, 1u);
        umma_commit(empty0 + 8 * s);
        if (kb == kKB - 1) umma_commit(acc0 + 8 * layer);        // this layer's accumulator is complete
      }
    }
    __syncwarp();
  } else {
    // ===== conversion + epilogue warps: TMEM lane quadrant = warp % 4, column half h = (warp - 2) / 4 =====
    const int quad = warp & 3, h = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = m < p.M;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    {
      const int e = threadIdx.x - 64;                  // 0..255
      s_vec[e] = p.b1[e];
      s_vec[C + e] = p.g1 ? p.g1[e] : 1.f;
      s_vec[2 * C + e] = p.g1 ? p.be1[e] : 0.f;
      s_vec[3 * C + e] = p.b2[e];
      s_vec[4 * C + e] = p.g2 ? p.g2[e] : 1.f;
      s_vec[5 * C + e] = p.g2 ? p.be2[e] : 0.f;
      if (e < N3P) s_b3[e] = e < p.N3 ? p.b3[e] : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }

    // ---- layers 1 and 2: accumulator -> A tiles of the next GEMM: f(acc + b) split into hi / lo, K-major SWIZZLE_128B
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
      const float* bias = s_vec + 3 * C * layer;
      const float* gam = bias + C;
      const float* bet = gam + C;
      const bool ln = (layer == 0 ? p.g1 : p.g2) != nullptr;
      const uint32_t tacc = tlane + (layer == 1 ? 256u : 0u);
      mbar_wait(acc0 + 8 * layer, 0);
      tc_fence_after();
      if (threadIdx.x == 64 && layer == 0) TC_TRACE(7);
      if (threadIdx.x == 64 && layer == 1) TC_TRACE(11);
      float mean = 0.f, rstd = 1.f;
      if (ln) {                            // statistics of the whole row: this thread's 128 columns + the other half's
        float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int j = 0; j < kKB; ++j) {
          uint32_t r[32];
          tmem_ld32(tacc + (uint32_t)(64 * j + 32 * h), r);
          const float* bj = bias + 64 * j + 32 * h;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float v = __uint_as_float(r[i]) + bj[i];
            p1[i & 3] += v; p2[i & 3] = fmaf(v, v, p2[i & 3]);
          }
        }
        const float s1 = (p1[0] + p1[1]) + (p1[2] + p1[3]), s2 = (p2[0] + p2[1]) + (p2[2] + p2[3]);
        float2* st = s_stat + layer * 2 * BM;
        st[h * BM + row] = make_float2(s1, s2);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float2 a = st[row], b = st[BM + row];           // summed in half order: identical statistics in both threads
        const float t1 = a.x + b.x, t2 = a.y + b.y;
        mean = t1 * (1.0f / C);
        rstd = rsqrtf(fmaxf(t2 * (1.0f / C) - mean * mean, 0.f) + p.ln_eps);
      }
#pragma unroll 1
      for (int j = 0; j < kKB; ++j) {
        const int it = (layer + 1) * kKB + j, s = it & 1;
        mbar_wait(empty0 + 8 * s, (uint32_t)((it - 2) >> 1) & 1);     // the MMAs that last read this stage have retired
        uint32_t r[32];
        tmem_ld32(tacc + (uint32_t)(64 * j + 32 * h), r);
        const int cb = 64 * j + 32 * h;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float v0 = __uint_as_float(r[2 * i]) + bias[cb + 2 * i], v1 = __uint_as_float(r[2 * i + 1]) + bias[cb + 2 * i + 1];
          if (ln) {
            v0 = fmaf((v0 - mean) * rstd, gam[cb + 2 * i], bet[cb + 2 * i]);
            v1 = fmaf((v1 - mean) * rstd, gam[cb + 2 * i + 1], bet[cb + 2 * i + 1]);
          }
          v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f);
          hi[i] = pack_bf16(v0, v1);
          lo[i] = pack_bf16(v0 - bf16_lo(hi[i]), v1 - bf16_hi(hi[i]));
        }
        const uint32_t rowa = smem_base + s * kStageBytes + (uint32_t)row * 128u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t off = (uint32_t)(((4 * h + c) ^ (row & 7)) << 4);
          sts128(rowa + off, hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
          sts128(rowa + kATile + off, lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
        }
        fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's (async proxy) reads
        __syncwarp();
        if (lane == 0) mbar_arrive(a20 + 8 * s);
      }
    }

    // ---- layer 3: [128, N3] result + bias, row-local tail, row stores (N3 = 10: 40-byte rows)
    mbar_wait(acc0 + 16, 0);
    tc_fence_after();
    if (threadIdx.x == 64) TC_TRACE(12);
    if (h == 0) {
      uint32_t r[32];
      tmem_ld32(tlane, r);
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + s_b3[i];
      if (row_ok) {
        if (p.tail.kind != TC_TAIL_NONE) apply_tail(p.tail, m, v);
        float* dst = p.out + (long long)m * p.ld_out;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < p.N3) dst[i] = v[i];
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
#ifdef TC_TRACE_BUILD
  if (trc && threadIdx.x == 0) { TC_TRACE(8); trc[9] = gtime(); }
#endif
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace tc

extern "C" int tc_mlp(const tc_mlp_args* a, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(a != nullptr, TC_ERR_NULL, "tc_mlp: null argument block");
  TC_REQUIRE(a->X && a->W1 && a->W2 && a->W3 && a->b1 && a->b2 && a->b3 && a->out_f32, TC_ERR_NULL,
             "tc_mlp: X, W1, b1, W2, b2, W3, b3 and out_f32 are required");
  TC_REQUIRE((a->ln1_gamma == nullptr) == (a->ln1_beta == nullptr) && (a->ln2_gamma == nullptr) == (a->ln2_beta == nullptr), TC_ERR_NULL,
             "tc_mlp: LayerNorm gamma and beta come in pairs");
  TC_REQUIRE(a->M > 0, TC_ERR_SHAPE, "tc_mlp: M must be positive (got %d)", a->M);
  TC_REQUIRE(a->C == C && a->N3 >= 1 && a->N3 <= N3P, TC_ERR_SHAPE,
             "tc_mlp: built for C = %d and N3 <= %d (got C = %d, N3 = %d): use three tc_linear calls", C, N3P, a->C, a->N3);
  TC_REQUIRE(a->ldx >= 2 * C && a->ldw1 >= 2 * C && a->ldw2 >= 2 * C && a->ldw3 >= 2 * C && a->ld_out_f32 >= a->N3, TC_ERR_SHAPE,
             "tc_mlp: split operands are [rows, 2C] (hi | lo)");
  TC_REQUIRE(al16(a->X) && al16(a->W1) && al16(a->W2) && al16(a->W3) && (a->ldx * 2) % 16 == 0 && (a->ldw1 * 2) % 16 == 0 &&
                 (a->ldw2 * 2) % 16 == 0 && (a->ldw3 * 2) % 16 == 0,
             TC_ERR_ALIGN, "tc_mlp: operands need 16-byte aligned bases and row pitches (TMA)");
  TC_REQUIRE(a->tail == TC_TAIL_NONE || (a->tail_in && a->N3 >= 8), TC_ERR_NULL, "tc_mlp: a tail needs tail_in and N3 >= 8");
  TC_REQUIRE(a->tail != TC_TAIL_REF_UPDATE || a->tail_ref_out, TC_ERR_NULL, "tc_mlp: the reference-update tail needs tail_ref_out");
  cudaStream_t s = as_stream(stream);
  CUtensorMap mx, mw1, mw2, mw3;
  if (!get_map(a->X, a->ldx, a->M, 2 * C, BM, &mx)) return TC_ERR_SHAPE;
  if (!get_map(a->W1, a->ldw1, C, 2 * C, 256, &mw1)) return TC_ERR_SHAPE;
  if (!get_map(a->W2, a->ldw2, C, 2 * C, 256, &mw2)) return TC_ERR_SHAPE;
  if (!get_map(a->W3, a->ldw3, a->N3, 2 * C, N3P, &mw3)) return TC_ERR_SHAPE;
  MlpParams p;
  p.M = a->M; p.N3 = a->N3;
  p.b1 = a->b1; p.g1 = a->ln1_gamma; p.be1 = a->ln1_beta;
  p.b2 = a->b2; p.g2 = a->ln2_gamma; p.be2 = a->ln2_beta;
  p.b3 = a->b3;
  p.ln_eps = a->ln_eps;
  p.out = a->out_f32; p.ld_out = a->ld_out_f32;
  p.tail.kind = a->tail; p.tail.in = a->tail_in; p.tail.ld_in = a->ld_tail_in;
  p.tail.ref_out = a->tail_ref_out; p.tail.geom_out = a->tail_geom_out;
  p.tail.xy_col = a->tail_xy_col; p.tail.z_col = a->tail_z_col; p.tail.from_norm = a->tail_from_norm;
  for (int i = 0; i < 6; ++i) p.tail.pc[i] = a->tail_pc_range[i];
  p.tail.r_lo = a->tail_r_lo; p.tail.r_hi = a->tail_r_hi;
  p.w_static = a->w_static ? 1 : 0;
  p.trace = trace_take((a->M + BM - 1) / BM);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) { set_error("tc_mlp: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  cudaError_t e = launch(mlp_tc_kernel, dim3((unsigned)((a->M + BM - 1) / BM)), dim3(kThreads), kSmemBytes, s, 1u, mx, mw1, mw2, mw3, p);
  if (e != cudaSuccess) { set_error("tc_mlp: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return (int)e; }
  count_launch();
  return check_launch("tc_mlp");
}
