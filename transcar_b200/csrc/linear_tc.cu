// K3 tensor-core path: Y = epilogue(A[M,K] * W[N,K]^T) on tcgen05, fp32 accumulation in TMEM.
//
// The GEMMs of this path are small (M = B*900 rows, N <= 1536, K <= 512): one 128 x 64 output tile per CTA keeps
// all 148 SMs busy (2 resident CTAs each) and makes the per-thread epilogue short.
//   * operands: TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) into a shared-memory ring, full/empty mbarriers; both
//     operands are K-major (nn.Linear keeps W as [N,K]).
//   * math: tcgen05.mma.cta_group::1.kind::f16, UMMA 128 x 64 x 16, issued by one thread; accumulator = 128 lanes x
//     64 fp32 columns of tensor memory.  Two operand modes:
//       TC_BF16    one pass over bf16 operands (4-stage ring, 24 KB per stage);
//       TC_BF16X2  "bf16x3": both operands arrive split (hi | lo halves of a [rows, 2K] bf16 matrix, tc_cast_split / the
//                  split epilogue below); per K block the ring stage holds A_hi, A_lo, W_hi, W_lo (48 KB, 2 stages) and
//                  the issuer runs hi*hi, lo*hi, hi*lo into the same accumulator: ~16 mantissa bits per operand, results
//                  within ~1e-5 of the fp32 reference for 3x the (otherwise idle) tensor-pipe work.
//   * the input-side epilogue terms (bias, per-query row bias, residuals) do not depend on the product: the four
//     epilogue warps request them while the operands are in flight, keep them in registers and add them after the
//     accumulator is complete (folding them into the accumulator BEFORE the MMAs made the MMAs wait for those loads).
//   * output side: one accumulator row per thread (tcgen05.ld 32x32b): LayerNorm / ReLU / post-add, 256-bit row
//     stores; fp32 and / or a 16-bit copy (bf16, split bf16 for the next bf16x3 GEMM, or fp16 for the attention core).
//     LayerNorm needs the whole row (N = 64, 128 or 256): the 1, 2 or 4 CTAs that share a row block form a
//     thread-block cluster and exchange per-row (sum, sum of squares) through distributed shared memory.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = epilogue (one
// accumulator row x 32 columns per thread).
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include <cuda_fp16.h>

#include "tc_common.cuh"
#include "tc_epilogue.cuh"
#include "tc_sm100.cuh"

namespace tc {
namespace {

constexpr int BM = 128;
constexpr int BN = 64;
constexpr int BK = 64;                 // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int kThreads = 320;               // producer warp, MMA warp, eight epilogue warps
constexpr int HN = 32;                      // columns per epilogue thread (two threads share an accumulator row)
constexpr uint32_t kABytes = BM * BK * 2;
constexpr uint32_t kBBytes = BN * BK * 2;
constexpr int kMaxCluster = 4;                        // LayerNorm rows span at most 4 CTAs (N <= 256)
constexpr int kSlots = 2 * kMaxCluster;               // statistics contributors per row: (CTA, column half)
// shared-memory carve-up per operand mode.  after the ring: barriers (full, empty, acc, init) + tmem slot, column
// vectors (bias, gamma, beta), LN partials
// (Measured and rejected: a 4-stage, 192 KB ring for split-mode grids of at most one CTA per SM - N <= 64, 57 CTAs at
// M = 7200.  5.0 vs ~6 us back to back, but inside the step such a CTA needs a whole SM to itself and waits behind the
// co-running branch's kernels: +30 us per step.)
template <bool kSplit>
struct Cfg {
  static constexpr int kStages = kSplit ? 2 : 4;
  static constexpr uint32_t kStageBytes = (kSplit ? 2u : 1u) * (kABytes + kBBytes);   // split: A_hi, A_lo, W_hi, W_lo
  static constexpr uint32_t kOffBars = kStages * kStageBytes;
  static constexpr uint32_t kOffVec = kOffBars + 128;
  static constexpr uint32_t kOffPart = kOffVec + 3 * BN * 4;
  static constexpr uint32_t kSmemUsed = kOffPart + 2 * kMaxCluster * BM * 8;       // LN partials: one slot per (CTA, column half)
  static constexpr size_t kSmemBytes = kSmemUsed + 1024;          // + alignment slack
};

struct EpiParams {
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
  int out16;      // TC_BF16, TC_BF16X2 (hi at column n, lo at column N + n) or TC_F16
  TailParams tail;
  int vec;        // 1: every row-wise operand is 32-byte aligned with a 32-byte multiple pitch -> 256-bit accesses
  int has_init;   // row_bias / residual(s) present
  int w_static;   // W is a parameter: its first tiles are requested before the dependency wait
  int tma_out;    // outputs leave through shared-memory staging + TMA tile stores (map_o32 / map_o16) instead of row stores
  unsigned long long* trace;   // tc_debug_trace: 16 uint64 per CTA, or null
};

template <bool kLN, bool kSplit, bool kTail>
__global__ void __launch_bounds__(kThreads, 2)
linear_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_o32, const __grid_constant__ CUtensorMap map_o16, const EpiParams p) {
  using C = Cfg<kSplit>;
  constexpr int kStages = C::kStages;
  constexpr uint32_t kStageBytes = C::kStageBytes, kOffBars = C::kOffBars, kOffVec = C::kOffVec, kOffPart = C::kOffPart;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // pointer arithmetic (not an integer round-trip) so the compiler keeps the shared address space: LDS/STS, not LD/ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2);
  float* s_bias = reinterpret_cast<float*>(smem + kOffVec);
  float* s_gamma = s_bias + BN;
  float* s_beta = s_gamma + BN;
  float2* s_part = reinterpret_cast<float2*>(smem + kOffPart);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kStages), accbar = smem_u32(bars + 2 * kStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int num_kb = (p.K + BK - 1) / BK;      // a partial last K block is zero-filled by TMA (both operands)
#ifdef TC_TRACE_BUILD
  unsigned long long* trc = p.trace ? p.trace + 16ull * (blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
  if (trc && threadIdx.x == 0) { trc[0] = gtime(); trc[10] = smid(); TC_TRACE(1); }
#endif
  pdl_trigger();

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(accbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Static weights (w_static): the W halves of the first ring stages are requested BEFORE the dependency wait - this CTA
  // is often resident while its predecessor kernel is still draining, and the weight tiles then arrive during that tail
  // instead of after it.  The stage's expect_tx covers both operands; the A tiles follow after the wait.
  const int num_pre = p.w_static ? min(num_kb, kStages) : 0;
  if (threadIdx.x == 0) TC_TRACE(2);
  if (threadIdx.x == 0) {
    for (int kb = 0; kb < num_pre; ++kb) {
      const uint32_t w_dst = smem_base + kb * kStageBytes + (kSplit ? 2 * kABytes : kABytes);
      mbar_expect_tx(full0 + 8 * kb, kStageBytes);
      tma_load_2d(w_dst, &map_w, kb * BK, n0, full0 + 8 * kb);
      if (kSplit) tma_load_2d(w_dst + kBBytes, &map_w, p.K + kb * BK, n0, full0 + 8 * kb);
    }
  }
  pdl_wait();            // everything above touched parameters only; activations are read from here on
  if (threadIdx.x == 0) TC_TRACE(3);
  if (kLN && warp < 2) cluster_arrive();   // these warps own no statistics slots: their share of the cluster barrier

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        const bool w_done = kb < num_pre;            // this stage's W tiles are already in flight (requested before the wait)
        if (!w_done) {
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          mbar_expect_tx(full0 + 8 * s, kStageBytes);
        }
        const uint32_t a_dst = smem_base + s * kStageBytes;
        if (kSplit) {            // stage = A_hi | A_lo | W_hi | W_lo; the lo halves start at column K of the [rows, 2K] matrices
          tma_load_2d(a_dst, &map_a, kb * BK, m0, full0 + 8 * s);
          if (!w_done) tma_load_2d(a_dst + 2 * kABytes, &map_w, kb * BK, n0, full0 + 8 * s);
          tma_load_2d(a_dst + kABytes, &map_a, p.K + kb * BK, m0, full0 + 8 * s);
          if (!w_done) tma_load_2d(a_dst + 2 * kABytes + kBBytes, &map_w, p.K + kb * BK, n0, full0 + 8 * s);
        } else {
          tma_load_2d(a_dst, &map_a, kb * BK, m0, full0 + 8 * s);
          if (!w_done) tma_load_2d(a_dst + kABytes, &map_w, kb * BK, n0, full0 + 8 * s);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // kind::f16 instruction descriptor: D=F32 (1<<4), A=BF16 (1<<7), B=BF16 (1<<10), K-major A and B,
      // N>>3 at bit 17, M>>4 at bit 24
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(full0 + 8 * s, ph);
        tc_fence_after();
        if (kb == 0) TC_TRACE(4);
        if (kb == num_kb - 1) TC_TRACE(5);
        const uint32_t a_addr = smem_base + s * kStageBytes;
        if (kSplit) {                              // hi*hi, then the two cross terms, into the same accumulator
          const uint64_t dah = make_desc_sw128(a_addr), dal = make_desc_sw128(a_addr + kABytes);
          const uint64_t dwh = make_desc_sw128(a_addr + 2 * kABytes), dwl = make_desc_sw128(a_addr + 2 * kABytes + kBBytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(tmem_base, dah + 2 * k, dwh + 2 * k, idesc, (kb | k) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_base, dal + 2 * k, dwh + 2 * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_base, dah + 2 * k, dwl + 2 * k, idesc, 1u);
        } else {
          const uint64_t da = make_desc_sw128(a_addr), db = make_desc_sw128(a_addr + kABytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)          // +32 bytes (2 x 16 B) per UMMA_K step inside the 128 B row
            umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(empty0 + 8 * s);               // frees the smem slot once these MMAs retire
      }
      umma_commit(accbar);                         // accumulator complete
    }
    __syncwarp();
  } else {
    // ===== epilogue: warps 2..9.  TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4: one accumulator row x 32
    // columns per thread.  (Four warps x 64 columns spent ~1000 cycles per 32-column chunk - ~220 dependent-issue-bound
    // instructions at ~4.5 cycles each with one or two warps per scheduler, tools/linear_trace.py; eight warps halve it.)
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = m < p.M;
    const uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * HN);
    const bool vec = p.vec != 0;
    const int c0 = half * HN;                                  // this thread's first column inside the tile
    const int nc = min(BN, p.N - n0) - c0;                     // its valid columns (<= 0: none)
    const int n = n0 + c0;
    const bool gate = (p.row_gate && row_ok) ? (p.row_gate[m] != 0) : true;
    if (kLN) {                                                 // statistics slots of this row: empty (each half clears its share)
      for (int r = half; r < kSlots; r += 2) reinterpret_cast<unsigned long long*>(s_part)[r * BM + row] = kStatSentinel;
      cluster_arrive();                                        // (warps 0 / 1 arrive before their roles start)
    }
    {                                                          // column vectors of this tile (zero beyond N)
      const int j = threadIdx.x - 64;
      if (j < BN) {
        const int nn = n0 + j;
        s_bias[j] = (p.bias && nn < p.N) ? p.bias[nn] : 0.f;
        s_gamma[j] = (kLN && nn < p.N) ? p.ln_gamma[nn] : 0.f;
        s_beta[j] = (kLN && nn < p.N) ? p.ln_beta[nn] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");          // the eight epilogue warps only
    }

    // Input-side terms (bias, per-query row bias, residuals) do not depend on the product.  Fast path (all 32 columns valid,
    // 32-byte aligned rows): they are requested right here, while the operands are in flight and the MMAs run, stay in
    // registers (32 per thread) and are added to the accumulator afterwards.  (Folding them into the accumulator BEFORE the
    // MMAs - the first version - made the MMAs wait for these loads: measured 4400 cycles for one fp32 residual tile vs 2100
    // for operands + MMAs, clock64.)  Partial tiles / unaligned rows fetch them after the accumulator is complete.
    const bool side_regs = p.has_init && vec && nc >= HN;
    float side[HN];
#pragma unroll
    for (int j = 0; j < HN; ++j) side[j] = 0.f;
    if (side_regs && row_ok) {
      auto add_side = [&](const float* src) {
        float t[HN];
#pragma unroll
        for (int i = 0; i < HN / 8; ++i) ld256(src + 8 * i, t + 8 * i);
#pragma unroll
        for (int j = 0; j < HN; ++j) side[j] += t[j];
      };
      if (gate) {
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < HN; ++j) side[j] = s_bias[c0 + j];
        }
        if (p.row_bias) add_side(p.row_bias + (long long)(m % p.row_bias_period) * p.ld_row_bias + n);
      }
      if (p.residual) add_side(p.residual + (long long)m * p.ld_residual + n);
      if (p.residual2) add_side(p.residual2 + (long long)m * p.ld_residual2 + n);
    }

    if (threadIdx.x == 64) TC_TRACE(11);
    mbar_wait(accbar, 0);
    tc_fence_after();
    if (threadIdx.x == 64) TC_TRACE(6);

    // pre-activation values of this thread:  (gate ? acc + bias + row_bias : 0) + residual + residual2
    float v[HN];
    {
      uint32_t r[HN];
      tmem_ld32(tcol, r);
      if (threadIdx.x == 64) TC_TRACE(12);
      if (side_regs) {                                     // gated-off rows: side[] holds the residuals only
#pragma unroll
        for (int j = 0; j < HN; ++j) v[j] = (gate ? __uint_as_float(r[j]) : 0.f) + side[j];
      } else {
#pragma unroll
        for (int j = 0; j < HN; ++j) v[j] = gate ? __uint_as_float(r[j]) : 0.f;
        if (gate && p.bias) {
#pragma unroll
          for (int j = 0; j < HN; ++j) v[j] += s_bias[c0 + j];
        }
        if (p.has_init && row_ok && nc > 0) {
          if (gate && p.row_bias) add_row32(p.row_bias + (long long)(m % p.row_bias_period) * p.ld_row_bias + n, nc, vec, v);
          if (p.residual) add_row32(p.residual + (long long)m * p.ld_residual + n, nc, vec, v);
          if (p.residual2) add_row32(p.residual2 + (long long)m * p.ld_residual2 + n, nc, vec, v);
        }
      }
    }
    // Output side.  Fast path (tma_out): every warp stages its 32 rows x 32 columns in shared memory - the operand ring is
    // free once the accumulator is complete - in the swizzled layout of the output tensor maps (conflict-free 128-bit
    // st.shared) and one lane issues TMA tile stores: whole 128 / 64-byte row segments per request instead of 32-byte
    // sectors of 32 different lines per st.global (the row stores held the epilogue at ~1 sector per clock and SM:
    // tools/linear_trace.py, 3700 of a CTA's 10400 cycles).  Rows beyond M / columns beyond N are clipped by the TMA unit.
    auto store_row = [&](float (&v)[HN]) {
      if (nc <= 0) return;
      if (p.post_add && row_ok) add_row32(p.post_add + (long long)m * p.ld_post_add + n, nc, vec, v);
      if (p.tma_out) {
        const uint32_t stg = smem_base + (uint32_t)(warp - 2) * 8192u;      // f32 tile 4 KB | 16-bit (hi) 2 KB | lo 2 KB
        if (p.out_f32) {
          const uint32_t rowa = stg + (uint32_t)lane * 128u;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            sts128(rowa + (uint32_t)((c ^ (lane & 7)) << 4), __float_as_uint(v[4 * c]), __float_as_uint(v[4 * c + 1]),
                   __float_as_uint(v[4 * c + 2]), __float_as_uint(v[4 * c + 3]));
        }
        if (p.out_bf16) {
          uint32_t u[16];
          if (p.out16 == TC_F16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) u[i] = pack_f16_sat(v[2 * i], v[2 * i + 1]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) u[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
          }
          const uint32_t rowa = stg + 4096u + (uint32_t)lane * 64u;
          const int sw = (lane >> 1) & 3;
#pragma unroll
          for (int c = 0; c < 4; ++c) sts128(rowa + (uint32_t)((c ^ sw) << 4), u[4 * c], u[4 * c + 1], u[4 * c + 2], u[4 * c + 3]);
          if (p.out16 == TC_BF16X2) {            // lo = bf16(v - hi) at column N + n
#pragma unroll
            for (int i = 0; i < 16; ++i) u[i] = pack_bf16(v[2 * i] - bf16_lo(u[i]), v[2 * i + 1] - bf16_hi(u[i]));
#pragma unroll
            for (int c = 0; c < 4; ++c) sts128(rowa + 2048u + (uint32_t)((c ^ sw) << 4), u[4 * c], u[4 * c + 1], u[4 * c + 2], u[4 * c + 3]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        const int mrow = m0 + quad * 32;
        if (lane == 0 && mrow < p.M) {
          if (p.out_f32) tma_store_2d(&map_o32, stg, n, mrow);
          if (p.out_bf16) {
            tma_store_2d(&map_o16, stg + 4096u, n, mrow);
            if (p.out16 == TC_BF16X2) tma_store_2d(&map_o16, stg + 6144u, p.N + n, mrow);
          }
          tma_store_commit();
        }
        return;
      }
      if (!row_ok) return;
      if (vec && nc >= HN) {
        if (p.out_f32) {
          float* dst = p.out_f32 + (long long)m * p.ld_out_f32 + n;
#pragma unroll
          for (int i = 0; i < 4; ++i) st256(dst + 8 * i, v + 8 * i);
        }
        if (p.out_bf16) {
          uint32_t u[16];
          __nv_bfloat16* dst = p.out_bf16 + (long long)m * p.ld_out_bf16 + n;
          if (p.out16 == TC_F16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) u[i] = pack_f16_sat(v[2 * i], v[2 * i + 1]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) u[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
          }
          st256u(dst, u);
          st256u(dst + 16, u + 8);
          if (p.out16 == TC_BF16X2) {            // lo = bf16(v - hi) at column N + n
#pragma unroll
            for (int i = 0; i < 16; ++i) u[i] = pack_bf16(v[2 * i] - bf16_lo(u[i]), v[2 * i + 1] - bf16_hi(u[i]));
            st256u(dst + p.N, u);
            st256u(dst + p.N + 16, u + 8);
          }
        }
      } else {
        if (p.out_f32) {
          float* dst = p.out_f32 + (long long)m * p.ld_out_f32 + n;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (vec && 8 * i + 8 <= nc) {
              st256(dst + 8 * i, v + 8 * i);
            } else {
#pragma unroll
              for (int j = 8 * i; j < 8 * i + 8; ++j)
                if (j < nc) dst[j] = v[j];
            }
          }
        }
        if (p.out_bf16) {
#pragma unroll
          for (int j = 0; j < HN; ++j) {
            if (j < nc) {
              __nv_bfloat16* d16 = p.out_bf16 + (long long)m * p.ld_out_bf16 + n + j;
              if (p.out16 == TC_F16) {
                *reinterpret_cast<__half*>(d16) = __float2half_rn(fminf(fmaxf(v[j], -65504.f), 65504.f));
              } else {
                const __nv_bfloat16 hi = __float2bfloat16_rn(v[j]);
                *d16 = hi;
                if (p.out16 == TC_BF16X2) d16[p.N] = __float2bfloat16_rn(v[j] - __bfloat162float(hi));
              }
            }
          }
        }
      }
    };

    if (!kLN) {
      if (nc > 0) {
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < HN; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (kTail && half == 0 && row_ok && n0 == 0) apply_tail(p.tail, m, v);
        store_row(v);
      }
    } else {
      // LayerNorm over the full row: this thread holds 32 of its N columns, its CTA 64, the cluster all of them.  Every
      // (CTA, half) pair is one contributor: 2 * cluster size partial sums per row, exchanged through shared memory.
      float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};   // four partial sums: 8-deep dependent chains, not 32
#pragma unroll
      for (int j = 0; j < HN; ++j) { p1[j & 3] += v[j]; p2[j & 3] = fmaf(v[j], v[j], p2[j & 3]); }
      float s1 = (p1[0] + p1[1]) + (p1[2] + p1[3]), s2 = (p2[0] + p2[1]) + (p2[2] + p2[3]);
      const uint32_t nct = cluster_nctarank();
      const uint32_t me = 2 * cluster_ctarank() + (uint32_t)half, nslot = 2 * nct;
      if (threadIdx.x == 64) TC_TRACE(13);
      cluster_wait();                                            // every peer has written its sentinels (long ago)
      {
        const unsigned long long mine = ((unsigned long long)__float_as_uint(s2) << 32) | __float_as_uint(s1);
        const uint32_t slot = smem_u32(&s_part[me * BM + row]);
        for (uint32_t r = 0; r < nct; ++r) st_dsmem_u64(slot, r, mine);     // into every CTA of the cluster, this one included
        float t1 = 0.f, t2 = 0.f;                                // summed in slot order: identical statistics in every thread
        for (uint32_t r = 0; r < nslot; ++r) {
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
      if (threadIdx.x == 64) TC_TRACE(14);
      const float inv_n = 1.0f / (float)p.N;
      const float mean = s1 * inv_n;
      const float var = fmaxf(s2 * inv_n - mean * mean, 0.f);    // biased variance, like nn.LayerNorm
      const float rstd = rsqrtf(var + p.ln_eps);
#pragma unroll
      for (int j = 0; j < HN; ++j) {
        v[j] = fmaf((v[j] - mean) * rstd, s_gamma[c0 + j], s_beta[c0 + j]);
        if (p.relu) v[j] = fmaxf(v[j], 0.f);
      }
      store_row(v);
    }
    if (p.tma_out && lane == 0) tma_store_wait_read();     // the staging tiles live in this CTA's shared memory
    tc_fence_before();
    if (threadIdx.x == 64) TC_TRACE(7);
#ifdef TC_TRACE_BUILD
    if (trc && lane == 0) atomicMax(trc + 15, (unsigned long long)clock64());     // the last epilogue warp to finish
#endif
  }
  if (kLN && warp < 2) cluster_wait();     // completes the arrive / wait pair of these warps
  // A CTA leaves only after it has received every peer's statistics, i.e. after the last write into its shared memory;
  // its own pushes target CTAs that are still waiting for them.
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
#ifdef TC_TRACE_BUILD
  if (trc && threadIdx.x == 0) { TC_TRACE(8); trc[9] = gtime(); }
#endif
}

// ---- host side: tensor maps -------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

struct MapKey {
  const void* ptr; long long ld; int rows, cols, box_rows, kind;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && ld == o.ld && rows == o.rows && cols == o.cols && box_rows == o.box_rows && kind == o.kind;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h = h * 1000003u ^ (size_t)k.ld; h = h * 1000003u ^ (size_t)k.rows; h = h * 1000003u ^ (size_t)k.cols;
    return h * 1000003u ^ (size_t)(k.box_rows * 4 + k.kind);
  }
};

}  // namespace

// see tc_epilogue.cuh
bool get_map(const void* ptr, long long ld, int rows, int cols, int box_rows, CUtensorMap* out, MapKind kind) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, ld, rows, cols, box_rows, (int)kind};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("tc_linear: cuTensorMapEncodeTiled entry point not available"); return false; }
  const bool f32 = kind == kMapOutF32;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (f32 ? 4 : 2)};
  cuuint32_t box[2] = {(cuuint32_t)(kind == kMapOperand ? BK : 32), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = fn(&m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, kind == kMapOut16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  kind == kMapOperand ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("tc_linear: cuTensorMapEncodeTiled failed (%d)", (int)r); return false; }
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return true;
}

namespace {

template <bool kLN, bool kSplit, bool kTail = false>
int launch_tile(const tc_linear_args* a, const EpiParams& ep, cudaStream_t s) {
  CUtensorMap ma, mw;
  const int kcols = kSplit ? 2 * a->K : a->K;       // split operands: [rows, 2K] = hi | lo
  constexpr size_t kSmemBytes = Cfg<kSplit>::kSmemBytes;
  if (!get_map(a->A, a->lda, a->M, kcols, BM, &ma)) return TC_ERR_SHAPE;
  if (!get_map(a->W, a->ldw, a->N, kcols, BN, &mw)) return TC_ERR_SHAPE;
  CUtensorMap mo32 = ma, mo16 = ma;                  // placeholders when an output (or the TMA-store path) is absent
  if (ep.tma_out) {
    if (a->out_f32 && !get_map(a->out_f32, a->ld_out_f32, a->M, a->N, 32, &mo32, kMapOutF32)) return TC_ERR_SHAPE;
    const int cols16 = ep.out16 == TC_BF16X2 ? 2 * a->N : a->N;
    if (a->out_bf16 && !get_map(a->out_bf16, a->ld_out_bf16, a->M, cols16, 32, &mo16, kMapOut16)) return TC_ERR_SHAPE;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(linear_tc_kernel<kLN, kSplit, kTail>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) { set_error("tc_linear: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  const unsigned ntiles = (unsigned)((a->N + BN - 1) / BN);
  // LayerNorm: the CTAs of one row block form a cluster and exchange row statistics
  cudaError_t e = launch(linear_tc_kernel<kLN, kSplit, kTail>, dim3(ntiles, (unsigned)((a->M + BM - 1) / BM)), dim3(kThreads), kSmemBytes, s,
                         kLN ? ntiles : 1u, ma, mw, mo32, mo16, ep);
  if (e != cudaSuccess) { set_error("tc_linear(tcgen05): %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return (int)e; }
  count_launch();
  return check_launch("tc_linear(tcgen05)");
}

inline bool al32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }
inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

// 256-bit row accesses: every row-wise operand 32-byte aligned, pitch a multiple of 32 bytes
static bool epilogue_vectorizable(const tc_linear_args* a) {
  if (a->row_bias && (!al32(a->row_bias) || a->ld_row_bias % 8 != 0)) return false;
  if (a->residual && (!al32(a->residual) || a->ld_residual % 8 != 0)) return false;
  if (a->residual2 && (!al32(a->residual2) || a->ld_residual2 % 8 != 0)) return false;
  if (a->post_add && (!al32(a->post_add) || a->ld_post_add % 8 != 0)) return false;
  if (a->out_f32 && (!al32(a->out_f32) || a->ld_out_f32 % 8 != 0)) return false;
  if (a->out_bf16 && (!al32(a->out_bf16) || a->ld_out_bf16 % 16 != 0)) return false;
  if (a->out_bf16 && a->out16_dtype == TC_BF16X2 && a->N % 16 != 0) return false;      // lo half starts at column N
  return true;
}

bool linear_tc_supported(const tc_linear_args* a) {
  static const bool disabled = getenv("TC_DISABLE_TC_LINEAR") != nullptr;        // debugging / A-B measurements
  if (disabled) return false;
  const bool split = a->a_dtype == TC_BF16X2 && a->w_dtype == TC_BF16X2;
  if (!split && (a->a_dtype != TC_BF16 || a->w_dtype != TC_BF16)) return false;
  if (a->K % 8 != 0 || a->K < BK) return false;       // 16-byte row pitch; a partial last K block is zero-filled
  if (split && a->K % BK != 0) return false;          // ... which the hi | lo layout does not allow
  // fused LayerNorm: the row block's CTAs form one cluster (portable size <= 8) -> N = 64, 128 or 256 here
  if (a->ln_gamma && !(a->N == 64 || a->N == 128 || a->N == 256)) return false;
  // TMA: 16-byte aligned base and row pitch for both operands
  if (!al16(a->A) || !al16(a->W) || (a->lda * 2) % 16 != 0 || (a->ldw * 2) % 16 != 0) return false;
  return true;
}

int linear_tc_launch(const tc_linear_args* a, cudaStream_t s) {
  EpiParams ep;
  ep.M = a->M; ep.N = a->N; ep.K = a->K;
  ep.bias = a->bias;
  ep.row_bias = a->row_bias; ep.row_bias_period = a->row_bias_period > 0 ? a->row_bias_period : 1;
  ep.ld_row_bias = a->ld_row_bias;
  ep.row_gate = a->row_gate;
  ep.residual = a->residual; ep.ld_residual = a->ld_residual;
  ep.residual2 = a->residual2; ep.ld_residual2 = a->ld_residual2;
  ep.ln_gamma = a->ln_gamma; ep.ln_beta = a->ln_beta; ep.ln_eps = a->ln_eps;
  ep.relu = a->relu;
  ep.post_add = a->post_add; ep.ld_post_add = a->ld_post_add;
  ep.out_f32 = a->out_f32; ep.ld_out_f32 = a->ld_out_f32;
  ep.out_bf16 = static_cast<__nv_bfloat16*>(a->out_bf16); ep.ld_out_bf16 = a->ld_out_bf16;
  ep.out16 = a->out16_dtype == 0 ? TC_BF16 : a->out16_dtype;
  ep.tail = make_tail(a);
  ep.vec = epilogue_vectorizable(a) ? 1 : 0;
  ep.has_init = (a->row_bias || a->residual || a->residual2) ? 1 : 0;
  static const bool no_prefetch = getenv("TC_NO_WPREFETCH") != nullptr;             // A/B measurements
  ep.w_static = (a->w_static && !no_prefetch) ? 1 : 0;
  // TMA-store epilogue: 16-byte aligned bases and row pitches; a split output needs whole 32-column tiles (a partial hi
  // tile would reach into the lo half)
  static const bool no_tma_store = getenv("TC_NO_TMA_STORE") != nullptr;             // A/B measurements
  ep.tma_out = !no_tma_store && (a->out_f32 || a->out_bf16) && a->N >= 64 &&     // (narrow outputs: the row stores are as fast)
               (!a->out_f32 || (al16(a->out_f32) && (a->ld_out_f32 * 4) % 16 == 0)) &&
               (!a->out_bf16 || (al16(a->out_bf16) && (a->ld_out_bf16 * 2) % 16 == 0 && (ep.out16 != TC_BF16X2 || a->N % 32 == 0)));
  ep.trace = trace_take((long long)((a->N + BN - 1) / BN) * ((a->M + BM - 1) / BM));
  if (a->a_dtype == TC_BF16X2) {
    if (a->tail) return launch_tile<false, true, true>(a, ep, s);
    return a->ln_gamma ? launch_tile<true, true>(a, ep, s) : launch_tile<false, true>(a, ep, s);
  }
  if (a->tail) return launch_tile<false, false, true>(a, ep, s);
  return a->ln_gamma ? launch_tile<true, false>(a, ep, s) : launch_tile<false, false>(a, ep, s);
}

}  // namespace tc

namespace tc {
namespace {
unsigned long long* g_trace_buf = nullptr;         // tc_debug_trace state: the buffer and the next free record
long long g_trace_cap = 0, g_trace_next = 0;
}  // namespace
unsigned long long* trace_take(long long ctas) {
  if (!g_trace_buf || g_trace_next + ctas > g_trace_cap) return nullptr;
  unsigned long long* r = g_trace_buf + 16 * g_trace_next;
  g_trace_next += ctas;
  return r;
}
}  // namespace tc

extern "C" int tc_debug_trace(uint64_t* device_buf, int64_t records) {
  tc::g_trace_buf = reinterpret_cast<unsigned long long*>(device_buf);
  tc::g_trace_cap = device_buf ? records : 0;
  tc::g_trace_next = 0;
  return TC_OK;
}
