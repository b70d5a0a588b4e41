// K3f fused feed-forward block on tcgen05:  Y = LayerNorm(X + W2 relu(W1 X + b1) + b2)  for C = 256, H = 512 in the bf16x3
// operand format (split bf16, three tensor-core passes per product, fp32 accumulation in tensor memory).
//
// Replaces two dependent tc_linear launches (mmcv FFN + norm of the decoder layers: 'ffn', 'norm' of the operation_order at
// projects/configs/detr3d/detr3d_res101_gridmask.py:65-82, and the radar head's rf_linear1 / rf_linear2 + rf_norm3, H:583-586 /
// H:655-658 / H:713-716).  Unfused, the [M, 512] hidden activation makes a round trip
// through L2 in split form (14.7 MB written, re-read once per 64-column output tile) and the second GEMM cannot start
// before the last tile of the first has left: 12.8 + 14.4 us back to back at M = 7200.  Here a 128-row block never
// leaves the SM pair that owns it:
//   * a cluster of two CTAs per 128 rows splits the HIDDEN dimension: CTA r computes hidden[:, 256 r : 256 r + 256]
//     (GEMM 1, N = 256, K = 256) into tensor memory columns 0..255, turns it - bias, ReLU, hi / lo split - 64 columns at a
//     time straight into the swizzled K-major A-operand tiles of GEMM 2 in shared memory, and accumulates its partial
//     product hidden_r W2[:, 256 r : 256 r + 256]^T (N = 256, K = 256) in columns 256..511;
//   * each CTA streams only its half of both weight matrices (512 KB in split form) through a 2 x 96 KB ring, the stage of
//     GEMM 2 being (A tile written by the conversion warps | W2 tile by TMA);
//   * the two partial sums meet through distributed shared memory: CTA r finalises output columns 128 r .. 128 r + 127; its
//     partner stages its partial for those columns in its own (by then idle) ring and sends the 64 KB as bulk copies
//     (cp.async.bulk shared::cta -> shared::cluster, completion counted on r's mbarrier) into r's ring - per-thread
//     st.shared::cluster moved the same bytes in 7100 cycles (~9 B / clock and direction, tools/ffn_trace.py); then residual +
//     bias + LayerNorm statistics exchange (the slot scheme of linear_tc.cu, four contributors per row) + TMA tile stores of
//     the fp32 and split outputs.
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
constexpr int C = 256;                  // model width: K of GEMM 1, N of GEMM 2
constexpr int HH = 256;                 // hidden columns per CTA (the pair splits H = 2 * HH)
constexpr int OWN = C / 2;              // output columns a CTA finalises
constexpr int kKB = C / BK;             // K blocks of GEMM 1 (= HH / BK K blocks of GEMM 2)
constexpr uint32_t kATile = BM * BK * 2;                       // 16 KB: 128 rows x 64 bf16
constexpr uint32_t kWTile = 256 * BK * 2;                      // 32 KB: 256 rows x 64 bf16
constexpr uint32_t kStageBytes = 2 * kATile + 2 * kWTile;      // A_hi | A_lo | W_hi | W_lo = 96 KB
constexpr int kStages = 2;
constexpr uint32_t kOffBars = kStages * kStageBytes;           // full[2], a2[2], empty[2], hbar, accbar, ringfree, xbar, ackbar, tmem slot
constexpr uint32_t kOffVec = kOffBars + 128;                   // b1 (HH), b2 (OWN), gamma (OWN), beta (OWN)
constexpr uint32_t kOffPart = kOffVec + (HH + 3 * OWN) * 4;    // LayerNorm partials: kSlots x BM x (sum, sum of squares)
constexpr int kSlots = 4;                                      // (CTA, column half) contributors per row
constexpr uint32_t kSmemUsed = kOffPart + kSlots * BM * 8;
constexpr size_t kSmemBytes = kSmemUsed + 1024;
// once every MMA of the pair has retired the ring is reused: the partner's partial sums, then the TMA-store staging tiles
constexpr uint32_t kXchgBytes = BM * OWN * 4;                  // 128 rows x 128 fp32 (16-byte chunks XOR-swizzled by row)
constexpr uint32_t kOffXchgIn = 0;                             // the partner's partial sums for this CTA's columns
constexpr uint32_t kOffXchgOut = 65536;                        // this CTA's partial sums for the partner's columns (source of the copy)
constexpr uint32_t kOffStage = 65536;                          // then: 8 warps x 2 chunks x (f32 4 KB | hi 2 KB | lo 2 KB)
static_assert(kOffStage + 16 * 8192 <= kStages * kStageBytes, "staging must fit in the ring");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

struct FfnParams {
  int M;
  const float* b1; const float* b2;
  const float* residual; long long ld_residual;
  const float* ln_gamma; const float* ln_beta; float ln_eps;
  int w_static;
  int has_o32, has_o16;
  unsigned long long* trace;   // tc_debug_trace: 16 uint64 per CTA, or null
};

__global__ void __launch_bounds__(kThreads, 1)
ffn_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
              const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_o32,
              const __grid_constant__ CUtensorMap map_o16, const FfnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  float* s_b1 = reinterpret_cast<float*>(smem + kOffVec);
  float* s_b2 = s_b1 + HH;
  float* s_gamma = s_b2 + OWN;
  float* s_beta = s_gamma + OWN;
  float2* s_part = reinterpret_cast<float2*>(smem + kOffPart);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), a20 = smem_u32(bars + 2), empty0 = smem_u32(bars + 4);
  const uint32_t hbar = smem_u32(bars + 6), accbar = smem_u32(bars + 7);
  const uint32_t ringfree = smem_u32(bars + 8), xbar = smem_u32(bars + 9), ackbar = smem_u32(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                  // which half of the hidden dimension / of the output columns
  const int m0 = blockIdx.y * BM;
#ifdef TC_TRACE_BUILD
  unsigned long long* trc = p.trace ? p.trace + 16ull * (blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
  if (trc && threadIdx.x == 0) { trc[0] = gtime(); trc[10] = smid(); TC_TRACE(1); }
#endif
  pdl_trigger();

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(a20 + 8 * s, 8); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(hbar, 1);
    mbar_init(accbar, 1);
    mbar_init(ringfree, 1);              // arrived by the PARTNER once its MMAs have retired (its copy may then land here ...
    mbar_init(xbar, 1);                  // ... and this one counts the bytes of that copy)
    mbar_init(ackbar, 1);                // arrived by the partner once THIS CTA's copy has landed there (its source is then idle)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(xbar, kXchgBytes);
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w2) : "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);        // hidden accumulator (256 columns) + output accumulator (256)
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
      tma_load_2d(w_dst, &map_w1, it * BK, (int)rank * HH, full0 + 8 * it);
      tma_load_2d(w_dst + kWTile, &map_w1, C + it * BK, (int)rank * HH, full0 + 8 * it);
    }
  }
  pdl_wait();            // everything above touched parameters only; activations are read from here on
  if (threadIdx.x == 0) TC_TRACE(3);
  // The one cluster barrier: both CTAs have initialised their mbarriers and statistics slots before any remote arrive / copy /
  // push.  Split: the epilogue warps arrive after their slot init, the producer / MMA warps once their loops are done (an
  // arrive.release in front of the first TMA request delayed it by ~1500 cycles), everybody waits right before the first
  // remote operation.

  if (warp == 0) {
    // ===== TMA producer: iterations 0..3 = GEMM 1 (X tile + W1 tile), 4..7 = GEMM 2 (W2 tile; its A tile is written by the
    // conversion warps) =====
    if (lane == 0) {
      for (int it = 0; it < 2 * kKB; ++it) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const uint32_t st = smem_base + s * kStageBytes, full = full0 + 8 * s;
        const bool w_done = it < num_pre;
        if (!w_done) {
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          mbar_expect_tx(full, it < kKB ? kStageBytes : 2 * kWTile);
        }
        if (it < kKB) {
          tma_load_2d(st, &map_x, it * BK, m0, full);
          tma_load_2d(st + kATile, &map_x, C + it * BK, m0, full);
          if (!w_done) {
            tma_load_2d(st + 2 * kATile, &map_w1, it * BK, (int)rank * HH, full);
            tma_load_2d(st + 2 * kATile + kWTile, &map_w1, C + it * BK, (int)rank * HH, full);
          }
        } else {
          const int j = it - kKB;                       // W2 is [C, 2H] = hi | lo; this CTA's K range starts at column rank * HH
          tma_load_2d(st + 2 * kATile, &map_w2, (int)rank * HH + j * BK, 0, full);
          tma_load_2d(st + 2 * kATile + kWTile, &map_w2, 2 * HH + (int)rank * HH + j * BK, 0, full);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer: UMMA 128 x 256 x 16, three passes (hi*hi, lo*hi, hi*lo) per K step =====
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int it = 0; it < 2 * kKB; ++it) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(full0 + 8 * s, ph);
        if (it >= kKB) mbar_wait(a20 + 8 * s, (uint32_t)((it - kKB) >> 1) & 1);
        tc_fence_after();
        if (it == 0) TC_TRACE(4);
        if (it == kKB - 1) TC_TRACE(5);
        if (it == kKB) TC_TRACE(2);
        if (it == 2 * kKB - 1) TC_TRACE(11);
        const uint32_t st = smem_base + s * kStageBytes;
        const uint32_t d = tmem_base + (it < kKB ? 0u : 256u);
        const uint64_t dah = make_desc_sw128(st), dal = make_desc_sw128(st + kATile);
        const uint64_t dwh = make_desc_sw128(st + 2 * kATile), dwl = make_desc_sw128(st + 2 * kATile + kWTile);
        const bool first = (it & (kKB - 1)) == 0;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(d, dah + 2 * k, dwh + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(d, dal + 2 * k, dwh + 2 * k, idesc, 1u);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(d, dah + 2 * k, dwl + 2 * k, idesc, 1u);
        umma_commit(empty0 + 8 * s);
        if (it == kKB - 1) umma_commit(hbar);          // hidden accumulator complete
      }
      umma_commit(accbar);                             // partial output accumulator complete
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
      s_b1[e] = p.b1[rank * HH + e];
      if (e < OWN) {
        s_b2[e] = p.b2[rank * OWN + e];
        s_gamma[e] = p.ln_gamma[rank * OWN + e];
        s_beta[e] = p.ln_beta[rank * OWN + e];
      }
      for (int r = h; r < kSlots; r += 2) reinterpret_cast<unsigned long long*>(s_part)[r * BM + row] = kStatSentinel;
      cluster_arrive();
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    // identity branch + output bias for this thread's 64 final columns: requested now, consumed after both GEMMs
    const int ncol0 = (int)rank * OWN + 64 * h;        // first of this thread's final output columns
    float res[64];
    if (row_ok) {
      const float* src = p.residual + (long long)m * p.ld_residual + ncol0;
#pragma unroll
      for (int i = 0; i < 8; ++i) ld256(src + 8 * i, res + 8 * i);
    } else {
#pragma unroll
      for (int j = 0; j < 64; ++j) res[j] = 0.f;
    }

    // ---- hidden accumulator -> A tiles of GEMM 2: relu(acc + b1) split into hi / lo, K-major SWIZZLE_128B layout
    mbar_wait(hbar, 0);
    tc_fence_after();
    if (threadIdx.x == 64) TC_TRACE(6);
#pragma unroll 1
    for (int j = 0; j < kKB; ++j) {
      const int s = j & 1;
      mbar_wait(empty0 + 8 * s, (uint32_t)((j + 2) >> 1) & 1);     // the MMAs that last read this stage have retired
      uint32_t r[32];
      tmem_ld32(tlane + (uint32_t)(64 * j + 32 * h), r);
      const float* bj = s_b1 + 64 * j + 32 * h;
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float v0 = fmaxf(__uint_as_float(r[2 * i]) + bj[2 * i], 0.f), v1 = fmaxf(__uint_as_float(r[2 * i + 1]) + bj[2 * i + 1], 0.f);
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
#pragma unroll
    for (int j = 0; j < 64; ++j) res[j] += s_b2[64 * h + j];

    // ---- partial sums meet: push this CTA's partial for the partner's columns into the partner's (idle) ring
    if (threadIdx.x == 64) TC_TRACE(7);
    mbar_wait(accbar, 0);
    tc_fence_after();
    if (threadIdx.x == 64) TC_TRACE(12);
    cluster_wait();                      // (completed long ago) the partner's mbarriers and slots are initialised
    const uint32_t partner = rank ^ 1u;
    if (threadIdx.x == 64) mbar_arrive_remote(ringfree, partner);       // this CTA's ring is idle: the partner's copy may land
    if (threadIdx.x == 64) TC_TRACE(13);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld32(tlane + 256u + partner * OWN + (uint32_t)(64 * h + 32 * c), r);
      const uint32_t rowa = smem_base + kOffXchgOut + (uint32_t)row * 512u;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        sts128(rowa + (uint32_t)(((16 * h + 8 * c + q) ^ (row & 7)) << 4), r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
    }
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (threadIdx.x == 64) {
      mbar_wait(ringfree, 0);            // the partner's MMAs have retired: its ring is idle
#pragma unroll
      for (int i = 0; i < 4; ++i)
        bulk_copy_to_cluster(smem_base + kOffXchgIn + i * (kXchgBytes / 4), smem_base + kOffXchgOut + i * (kXchgBytes / 4),
                             kXchgBytes / 4, xbar, partner);
    }
    mbar_wait(xbar, 0);                  // the partner's partial sums have landed here
    if (threadIdx.x == 64) mbar_arrive_remote(ackbar, partner);
    if (threadIdx.x == 64) TC_TRACE(14);

    // ---- final rows: own partial + partner's + (bias + identity), LayerNorm over the 256 columns of the pair
    float v[64];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld32(tlane + 256u + rank * OWN + (uint32_t)(64 * h + 32 * c), r);
      const uint32_t rowa = smem_base + kOffXchgIn + (uint32_t)row * 512u;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float o[4];
        lds128(rowa + (uint32_t)(((16 * h + 8 * c + q) ^ (row & 7)) << 4), o);
#pragma unroll
        for (int t = 0; t < 4; ++t) v[32 * c + 4 * q + t] = (__uint_as_float(r[4 * q + t]) + o[t]) + res[32 * c + 4 * q + t];
      }
    }
    float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 64; ++j) { p1[j & 3] += v[j]; p2[j & 3] = fmaf(v[j], v[j], p2[j & 3]); }
    float s1 = (p1[0] + p1[1]) + (p1[2] + p1[3]), s2 = (p2[0] + p2[1]) + (p2[2] + p2[3]);
    {
      const uint32_t me = 2 * rank + (uint32_t)h;
      const unsigned long long mine = ((unsigned long long)__float_as_uint(s2) << 32) | __float_as_uint(s1);
      const uint32_t slot = smem_u32(&s_part[me * BM + row]);
      st_dsmem_u64(slot, 0, mine);
      st_dsmem_u64(slot, 1, mine);
      float t1 = 0.f, t2 = 0.f;                                // summed in slot order: identical statistics in every thread
      for (uint32_t r = 0; r < (uint32_t)kSlots; ++r) {
        float o1 = s1, o2 = s2;
        if (r != me) {
          const uint32_t addr = smem_u32(&s_part[r * BM + row]);
          unsigned long long w = ld_smem_u64(addr);
          for (uint32_t spins = 0; w == kStatSentinel && spins < kSpinLimit; ++spins) w = ld_smem_u64(addr);
          o1 = __uint_as_float((uint32_t)w); o2 = __uint_as_float((uint32_t)(w >> 32));
        }
        t1 += o1; t2 += o2;
      }
      s1 = t1; s2 = t2;
    }
    const float inv_n = 1.0f / (float)C;
    const float mean = s1 * inv_n;
    const float var = fmaxf(s2 * inv_n - mean * mean, 0.f);    // biased variance, like nn.LayerNorm
    const float rstd = rsqrtf(var + p.ln_eps);
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = fmaf((v[j] - mean) * rstd, s_gamma[64 * h + j], s_beta[64 * h + j]);

    if (threadIdx.x == 64) TC_TRACE(15);
    // ---- outputs: swizzled staging tiles + TMA tile stores (rows beyond M are clipped), as in linear_tc.cu.  The staging
    // tiles reuse the source of the outgoing copy: the partner acknowledges its arrival (a copy that completes on an
    // mbarrier is not part of a bulk async-group, so wait_group cannot vouch for it); the wait is over long before
    if (threadIdx.x == 64) mbar_wait(ackbar, 0);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int mrow = m0 + quad * 32;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint32_t stg = smem_base + kOffStage + (uint32_t)((warp - 2) * 2 + c) * 8192u;
      const float* vc = v + 32 * c;
      if (p.has_o32) {
        const uint32_t rowa = stg + (uint32_t)lane * 128u;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          sts128(rowa + (uint32_t)((q ^ (lane & 7)) << 4), __float_as_uint(vc[4 * q]), __float_as_uint(vc[4 * q + 1]),
                 __float_as_uint(vc[4 * q + 2]), __float_as_uint(vc[4 * q + 3]));
      }
      if (p.has_o16) {
        uint32_t u[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) u[i] = pack_bf16(vc[2 * i], vc[2 * i + 1]);
        const uint32_t rowa = stg + 4096u + (uint32_t)lane * 64u;
        const int sw = (lane >> 1) & 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) sts128(rowa + (uint32_t)((q ^ sw) << 4), u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
#pragma unroll
        for (int i = 0; i < 16; ++i) u[i] = pack_bf16(vc[2 * i] - bf16_lo(u[i]), vc[2 * i + 1] - bf16_hi(u[i]));
#pragma unroll
        for (int q = 0; q < 4; ++q) sts128(rowa + 2048u + (uint32_t)((q ^ sw) << 4), u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && mrow < p.M) {
        const int n = ncol0 + 32 * c;
        if (p.has_o32) tma_store_2d(&map_o32, stg, n, mrow);
        if (p.has_o16) {
          tma_store_2d(&map_o16, stg + 4096u, n, mrow);
          tma_store_2d(&map_o16, stg + 6144u, C + n, mrow);
        }
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait_read();
    tc_fence_before();
  }
  if (warp < 2) { cluster_arrive(); cluster_wait(); }       // these warps' share of the cluster barrier
  // A CTA leaves only after it has received the partner's partial sums and statistics, i.e. after the last write into its
  // shared memory; its own copy / pushes target a CTA that is still waiting for them.
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
inline bool al32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

}  // namespace
}  // namespace tc

extern "C" int tc_ffn(const tc_ffn_args* a, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(a != nullptr, TC_ERR_NULL, "tc_ffn: null argument block");
  TC_REQUIRE(a->X && a->W1 && a->W2 && a->b1 && a->b2 && a->residual && a->ln_gamma && a->ln_beta, TC_ERR_NULL,
             "tc_ffn: X, W1, b1, W2, b2, residual, ln_gamma and ln_beta are required");
  TC_REQUIRE(a->out_f32 || a->out16, TC_ERR_NULL, "tc_ffn: no output requested");
  TC_REQUIRE(a->M > 0, TC_ERR_SHAPE, "tc_ffn: M must be positive (got %d)", a->M);
  TC_REQUIRE(a->C == C && a->H == 2 * HH, TC_ERR_SHAPE, "tc_ffn: built for C = %d, H = %d (got C = %d, H = %d): use two tc_linear calls",
             C, 2 * HH, a->C, a->H);
  TC_REQUIRE(a->ldx >= 2 * C && a->ldw1 >= 2 * C && a->ldw2 >= 4 * HH, TC_ERR_SHAPE, "tc_ffn: split operands are [rows, 2K] (hi | lo)");
  TC_REQUIRE(al16(a->X) && al16(a->W1) && al16(a->W2) && (a->ldx * 2) % 16 == 0 && (a->ldw1 * 2) % 16 == 0 && (a->ldw2 * 2) % 16 == 0,
             TC_ERR_ALIGN, "tc_ffn: operands need 16-byte aligned bases and row pitches (TMA)");
  TC_REQUIRE(al32(a->residual) && a->ld_residual % 8 == 0 && a->ld_residual >= C, TC_ERR_ALIGN,
             "tc_ffn: residual rows must be 32-byte aligned");
  TC_REQUIRE(!a->out_f32 || (al16(a->out_f32) && (a->ld_out_f32 * 4) % 16 == 0 && a->ld_out_f32 >= C), TC_ERR_ALIGN,
             "tc_ffn: out_f32 needs a 16-byte aligned base and row pitch");
  TC_REQUIRE(!a->out16 || (al16(a->out16) && (a->ld_out16 * 2) % 16 == 0 && a->ld_out16 >= 2 * C), TC_ERR_ALIGN,
             "tc_ffn: out16 is split bf16 [M, 2C] with a 16-byte aligned base and row pitch");
  cudaStream_t s = as_stream(stream);
  CUtensorMap mx, mw1, mw2, mo32, mo16;
  if (!get_map(a->X, a->ldx, a->M, 2 * C, BM, &mx)) return TC_ERR_SHAPE;
  if (!get_map(a->W1, a->ldw1, 2 * HH, 2 * C, 256, &mw1)) return TC_ERR_SHAPE;
  if (!get_map(a->W2, a->ldw2, C, 4 * HH, 256, &mw2)) return TC_ERR_SHAPE;
  mo32 = mx; mo16 = mx;
  if (a->out_f32 && !get_map(a->out_f32, a->ld_out_f32, a->M, C, 32, &mo32, kMapOutF32)) return TC_ERR_SHAPE;
  if (a->out16 && !get_map(a->out16, a->ld_out16, a->M, 2 * C, 32, &mo16, kMapOut16)) return TC_ERR_SHAPE;
  FfnParams p;
  p.M = a->M; p.b1 = a->b1; p.b2 = a->b2;
  p.residual = a->residual; p.ld_residual = a->ld_residual;
  p.ln_gamma = a->ln_gamma; p.ln_beta = a->ln_beta; p.ln_eps = a->ln_eps;
  p.w_static = a->w_static ? 1 : 0;
  p.has_o32 = a->out_f32 ? 1 : 0; p.has_o16 = a->out16 ? 1 : 0;
  p.trace = trace_take(2ll * ((a->M + BM - 1) / BM));
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ffn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) { set_error("tc_ffn: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  cudaError_t e = launch(ffn_tc_kernel, dim3(2, (unsigned)((a->M + BM - 1) / BM)), dim3(kThreads), kSmemBytes, s, 2u,
                         mx, mw1, mw2, mo32, mo16, p);
  if (e != cudaSuccess) { set_error("tc_ffn: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return (int)e; }
  count_launch();
  return check_launch("tc_ffn");
}
