// K3c: row-local Linear chain - up to 12 dependent Linear layers for one 128-row tile per CTA, every intermediate on chip.
// Contract and the reference lines it replaces: include/transcar_b200.h (tc_linear_chain).
//
// Why: the decoder-layer tail (output_proj+LN, FFN+LN, 3 regression layers, reference update) is 7 dependent launches of
// <= 0.94 GFLOP whose operands sit in L2; as separate kernels each pays launch + TMA fill + TMEM drain (10-18 us each,
// profiles/r01_step_timeline.txt) for < 1 us of tensor work.  All of it is row-local, so one CTA can carry 128 rows through
// the whole chain:
//   * activations: two 64 KB shared-memory buffers (X, H), each [128 rows x 256 K] bf16 stored as four K-major
//     SWIZZLE_128B blocks - exactly the layout TMA produces and tcgen05.mma consumes, written by the epilogue threads;
//   * weights: streamed by TMA through a 3 x 32 KB ring ([<=256 rows x 64 K] per slot) by a producer warp that runs ahead
//     of the stage boundaries (the next layer's weights arrive while the current epilogue runs); weights are constants, so
//     the stream starts before griddepcontrol.wait;
//   * math: tcgen05.mma kind::f16 128 x N x 16 (N <= 256), one issuing thread, fp32 accumulators in tensor memory (all 512
//     columns are allocated: columns are named by the program, e.g. 0-255 = residual stream, 256-511 = hidden scratch);
//   * residuals never leave tensor memory: a LayerNorm epilogue writes its fp32 result (+ the bias of the layer that will
//     add onto it) back into accumulator columns and the later layer's MMAs accumulate on top (FFN: x + W2 relu(W1 x));
//   * epilogues: one accumulator row per thread (tcgen05.ld 32x32b), so LayerNorm needs no cross-thread reduction.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = init / epilogue.
// Per-stage single-use mbarriers: ready[s] (128 epilogue threads: operand buffer written, accumulator pre-loaded) and
// done[s] (tcgen05.commit: the stage's MMAs - and all earlier ones - have retired).
#include <cuda.h>

#include "tc_common.cuh"
#include "tc_sm100.cuh"

namespace tc {

bool tensor_map_bf16_2d(const void* ptr, long long ld, int rows, int cols, int box_rows, CUtensorMap* out);   // linear_tc.cu

namespace {

constexpr int CM = 128;                          // rows per CTA
constexpr int CK = 64;                           // K block: 64 bf16 = one 128-byte swizzle row
constexpr int kWStages = 3;
constexpr int kEpiWarps = 8;                     // two per TMEM lane quadrant: each handles half of a row's columns
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 64 + kEpiThreads;
constexpr int kMaxStages = TC_CHAIN_MAX_STAGES;
constexpr uint32_t kKBlockBytes = CM * 128;      // one K block of an activation buffer (16 KB)
constexpr uint32_t kActBytes = 4 * kKBlockBytes; // K <= 256
constexpr uint32_t kWSlotBytes = 256 * 128;      // one K block of a weight: <= 256 rows x 128 B
constexpr uint32_t kOffX = 0;
constexpr uint32_t kOffH = kActBytes;
constexpr uint32_t kOffW = 2 * kActBytes;
constexpr uint32_t kOffBars = kOffW + kWStages * kWSlotBytes;
// barriers: full[3], empty[3], abar, ready[12], done[12]  -> 31 x 8 bytes, then the TMEM slot
constexpr uint32_t kNumBars = 2 * kWStages + 1 + 2 * kMaxStages;
constexpr uint32_t kOffPart = kOffBars + kNumBars * 8 + 16;       // LayerNorm partial sums: [2 halves][128 rows] float2
constexpr uint32_t kSmemUsed = kOffPart + 2 * CM * 8;
// The buffers fill the SM's shared memory to within 1 KB, so there is no room for an alignment pad: the dynamic
// shared-memory window of a kernel without static shared memory starts 1024-byte aligned; the kernel traps otherwise.
constexpr size_t kSmemBytes = kSmemUsed;
static_assert(kSmemBytes <= 232448, "chain kernel: shared memory budget");

struct ChainParams {
  CUtensorMap map_a;
  CUtensorMap map_w[kMaxStages];
  int M, K0, nstages;
  tc_chain_stage st[kMaxStages];
};

__device__ __forceinline__ void ld256f(const float* p, float* v) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void st256f(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void st256b(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// v[j] += vec[j], j < 32: a column vector (bias / gamma / beta); every lane reads the same address (broadcast)
__device__ __forceinline__ void add_vec32(const float* vec, float (&v)[32]) {
  const float4* p = reinterpret_cast<const float4*>(vec);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = __ldg(p + i);
    v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
  }
}
__device__ __forceinline__ void load_vec32(const float* vec, float (&v)[32]) {
  const float4* p = reinterpret_cast<const float4*>(vec);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = __ldg(p + i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
// v[j] += row[j], j < 32: 32 consecutive floats of this thread's row (4 x 256-bit loads)
__device__ __forceinline__ void add_row32(const float* row, float (&v)[32]) {
  float t[32];
#pragma unroll
  for (int i = 0; i < 4; ++i) ld256f(row + 8 * i, t + 8 * i);
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] += t[j];
}
// 32 columns (chunk c) of this thread's row -> bf16 in a K-major SWIZZLE_128B activation buffer
__device__ __forceinline__ void store_act32(uint8_t* buf, int row, int c, const float (&v)[32]) {
  uint8_t* base = buf + (c >> 1) * kKBlockBytes + row * 128;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 u;
    u.x = pack_bf16(v[8 * g + 0], v[8 * g + 1]); u.y = pack_bf16(v[8 * g + 2], v[8 * g + 3]);
    u.z = pack_bf16(v[8 * g + 4], v[8 * g + 5]); u.w = pack_bf16(v[8 * g + 6], v[8 * g + 7]);
    const int chunk = (c & 1) * 4 + g;
    *reinterpret_cast<uint4*>(base + ((chunk ^ (row & 7)) << 4)) = u;
  }
}

// two 32-column accumulator chunks in flight, one wait
__device__ __forceinline__ void tmem_ld32x2(uint32_t ta, uint32_t tb, uint32_t (&a)[32], uint32_t (&b)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%64];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%65];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
        "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]),
        "=r"(a[16]), "=r"(a[17]), "=r"(a[18]), "=r"(a[19]), "=r"(a[20]), "=r"(a[21]), "=r"(a[22]), "=r"(a[23]),
        "=r"(a[24]), "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]), "=r"(a[30]), "=r"(a[31]),
        "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]),
        "=r"(b[8]), "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15]),
        "=r"(b[16]), "=r"(b[17]), "=r"(b[18]), "=r"(b[19]), "=r"(b[20]), "=r"(b[21]), "=r"(b[22]), "=r"(b[23]),
        "=r"(b[24]), "=r"(b[25]), "=r"(b[26]), "=r"(b[27]), "=r"(b[28]), "=r"(b[29]), "=r"(b[30]), "=r"(b[31])
      : "r"(ta), "r"(tb)
      : "memory");
}
__device__ __forceinline__ void tmem_st32_nowait(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__global__ void __launch_bounds__(kThreads, 1) chain_kernel(const __grid_constant__ ChainParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();              // SWIZZLE_128B tiles need 1024-byte alignment (see kSmemBytes)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t full0 = bar0, empty0 = bar0 + 8 * kWStages, abar = bar0 + 16 * kWStages,
                 ready0 = abar + 8, done0 = ready0 + 8 * kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  float2* s_part = reinterpret_cast<float2*>(smem + kOffPart);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * CM;
  const int S = P.nstages;
  pdl_trigger();

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(abar, 1);
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(ready0 + 8 * s, kEpiThreads); mbar_init(done0 + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_a) : "memory");
    for (int s = 0; s < S; ++s) asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_w[s]) : "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: weights of every stage through the ring; the activation tile once =====
    if (lane == 0) {
      bool a_issued = false;
      auto issue_a = [&]() {
        pdl_wait();                                   // the A tile is produced by the previous kernel
        const int nkb = P.K0 / CK;
        const uint32_t dst = sbase + (P.st[0].a_buf ? kOffH : kOffX);
        mbar_expect_tx(abar, (uint32_t)nkb * kKBlockBytes);
        for (int kb = 0; kb < nkb; ++kb) tma_load_2d(dst + kb * kKBlockBytes, &P.map_a, kb * CK, m0, abar);
        a_issued = true;
      };
      int it = 0;
      for (int s = 0; s < S; ++s) {
        const int nkb = P.st[s].K / CK;
        const uint32_t bytes = (uint32_t)((P.st[s].N + 15) & ~15) * 128u;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          if (it == kWStages && !a_issued) issue_a();   // before the first wait that needs the MMA warp to make progress
          const int slot = it % kWStages;
          mbar_wait(empty0 + 8 * slot, ((it / kWStages) & 1) ^ 1);
          mbar_expect_tx(full0 + 8 * slot, bytes);
          tma_load_2d(sbase + kOffW + slot * kWSlotBytes, &P.map_w[s], kb * CK, 0, full0 + 8 * slot);
        }
      }
      if (!a_issued) issue_a();
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int it = 0;
      for (int s = 0; s < S; ++s) {
        const int nkb = P.st[s].K / CK;
        const uint32_t n_mma = (uint32_t)((P.st[s].N + 15) & ~15);
        // kind::f16: D=F32 (1<<4), A=BF16 (1<<7), B=BF16 (1<<10), K-major A and B, N>>3 at bit 17, M>>4 at bit 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((n_mma >> 3) << 17) | ((uint32_t)(CM >> 4) << 24);
        const uint32_t a_base = sbase + (P.st[s].a_buf ? kOffH : kOffX);
        const uint32_t d_addr = tmem_base + (uint32_t)P.st[s].acc_col;
        const uint32_t acc0 = P.st[s].accumulate ? 1u : 0u;
        mbar_wait(ready0 + 8 * s, 0);                 // operand buffer written / accumulator pre-loaded
        if (s == 0) mbar_wait(abar, 0);
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int slot = it % kWStages;
          mbar_wait(full0 + 8 * slot, (it / kWStages) & 1);
          tc_fence_after();
          const uint64_t da = make_desc_sw128(a_base + kb * kKBlockBytes);
          const uint64_t db = make_desc_sw128(sbase + kOffW + slot * kWSlotBytes);
#pragma unroll
          for (int k = 0; k < CK / 16; ++k)
            umma_bf16(d_addr, da + 2 * k, db + 2 * k, idesc, (acc0 | (uint32_t)kb | (uint32_t)k) ? 1u : 0u);
          umma_commit(empty0 + 8 * slot);
        }
        umma_commit(done0 + 8 * s);
      }
    }
    __syncwarp();
  } else {
    // ===== init / epilogue: warps 2..9.  TMEM lane quadrant = warp % 4 (hardware rule); the two warps of a quadrant
    // split a row's 32-column chunks in halves.  One row per thread. =====
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = m < P.M;
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int etid = threadIdx.x - 64;
    uint8_t* bufX = smem + kOffX;
    uint8_t* bufH = smem + kOffH;
    pdl_wait();                                      // residuals / gates come from earlier kernels; outputs are written below

#pragma unroll 1
    for (int s = 0; s < S; ++s) {
      const tc_chain_stage& st = P.st[s];
      const int nchunks = (st.N + 31) >> 5;
      const int c_mid = (nchunks + 1) >> 1;
      const int c_begin = half ? c_mid : 0, c_end = half ? nchunks : c_mid;
      // column vectors of this stage -> L1 (they are read chunk by chunk below, by every thread at the same address)
      if (etid < nchunks) {
        if (st.bias) prefetch_l1(st.bias + etid * 32);
        if (st.ln_gamma) { prefetch_l1(st.ln_gamma + etid * 32); prefetch_l1(st.ln_beta + etid * 32); }
        if (st.fold_bias) prefetch_l1(st.fold_bias + etid * 32);
      }
      // ---- accumulator pre-load: (gate ? init_bias : 0) + residual + residual2 -------------------------------
      if (st.init) {
        const bool gate = (st.row_gate && row_ok) ? (st.row_gate[m] != 0) : true;
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 2) {
          const bool two = c + 1 < c_end;
          float v[2][32];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[h][j] = 0.f;
          }
          if (row_ok) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (h == 0 || two) {
                const int col = (c + h) * 32;
                if (st.init_bias && gate) load_vec32(st.init_bias + col, v[h]);
                if (st.residual) add_row32(st.residual + (long long)m * st.ld_residual + col, v[h]);
                if (st.residual2) add_row32(st.residual2 + (long long)m * st.ld_residual2 + col, v[h]);
              }
            }
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 0 || two) {
              uint32_t r[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(v[h][j]);
              tmem_st32_nowait(trow + (uint32_t)st.acc_col + (c + h) * 32, r);
            }
          }
        }
        tmem_wait_st();
      }
      fence_proxy_async_smem();        // this thread's activation-buffer writes (previous epilogue) -> async proxy
      tc_fence_before();               // this thread's tcgen05.st / tcgen05.ld are ordered before the MMAs
      mbar_arrive(ready0 + 8 * s);
      if (st.epi == TC_CHAIN_NONE) continue;

      mbar_wait(done0 + 8 * s, 0);
      tc_fence_after();
      const uint32_t tacc = trow + (uint32_t)st.acc_col;
      uint8_t* dst = st.dst_buf < 0 ? nullptr : (st.dst_buf ? bufH : bufX);

      if (st.epi == TC_CHAIN_OUT) {
        // ---- small output (N <= 32): bias, per-query row bias, optional pointwise tail; half 0 only ----------------
        if (half != 0) continue;
        uint32_t r[32];
        tmem_ld32(tacc, r);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (row_ok) {
          const int n = st.N;
          if (st.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < n) v[j] += __ldg(st.bias + j);
          }
          if (st.row_bias) {
            const float* rb = st.row_bias + (long long)(m % st.row_bias_period) * st.ld_row_bias;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < n) v[j] += rb[j];
          }
          if (st.tail == TC_CHAIN_TAIL_ANCHOR_ADD) {
            const float* an = st.tail_in + (long long)m * st.ld_tail_in;
            float ax = an[st.tail_xy_col], ay = an[st.tail_xy_col + 1];
            const float az = an[st.tail_z_col];
            if (st.tail_from_norm) {
              ax = __fadd_rn(__fmul_rn(ax, st.pc_range[3] - st.pc_range[0]), st.pc_range[0]);
              ay = __fadd_rn(__fmul_rn(ay, st.pc_range[4] - st.pc_range[1]), st.pc_range[1]);
            }
            v[0] = __fadd_rn(v[0], ax);
            v[1] = __fadd_rn(v[1], ay);
            v[4] = __fadd_rn(v[4], az);
          }
          if (st.out_f32) {
            float* o = st.out_f32 + (long long)m * st.ld_out_f32;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < n) o[j] = v[j];
          }
          if (st.tail == TC_CHAIN_TAIL_REF_UPDATE) {
            const float* rin = st.tail_in + (long long)m * st.ld_tail_in;
            float* ro = st.tail_out + (long long)m * 3;
            ro[0] = sigmoid_f32(__fadd_rn(v[0], logit_f32(rin[0])));
            ro[1] = sigmoid_f32(__fadd_rn(v[1], logit_f32(rin[1])));
            ro[2] = sigmoid_f32(__fadd_rn(v[4], logit_f32(rin[2])));
          }
        }
        continue;
      }

      // ---- wide epilogues (N % 32 == 0): this thread owns chunks [c_begin, c_end) of its row ------------------------
      float mean = 0.f, rstd = 1.f;
      const bool ln = st.epi == TC_CHAIN_LN;
      if (ln) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 2) {
          const bool two = c + 1 < c_end;
          uint32_t r[2][32];
          if (two) tmem_ld32x2(tacc + c * 32, tacc + (c + 1) * 32, r[0], r[1]);
          else tmem_ld32(tacc + c * 32, r[0]);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 0 || two) {
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[h][j]);
              if (st.bias) add_vec32(st.bias + (c + h) * 32, v);
#pragma unroll
              for (int j = 0; j < 32; ++j) { s1 += v[j]; s2 = fmaf(v[j], v[j], s2); }
            }
          }
        }
        s_part[half * CM + row] = make_float2(s1, s2);
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");      // the eight epilogue warps only
        const float2 o = s_part[(half ^ 1) * CM + row];
        s1 += o.x; s2 += o.y;
        const float inv_n = 1.0f / (float)st.N;
        mean = s1 * inv_n;
        const float var = fmaxf(s2 * inv_n - mean * mean, 0.f);        // biased variance, like nn.LayerNorm
        rstd = rsqrtf(var + st.ln_eps);
      }
#pragma unroll 1
      for (int c = c_begin; c < c_end; c += 2) {
        const bool two = c + 1 < c_end;
        uint32_t r[2][32];
        if (two) tmem_ld32x2(tacc + c * 32, tacc + (c + 1) * 32, r[0], r[1]);
        else tmem_ld32(tacc + c * 32, r[0]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 0 || two) {
            const int cc = c + h;
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[h][j]);
            if (st.bias) add_vec32(st.bias + cc * 32, v);
            if (ln) {
              float g[32];
              load_vec32(st.ln_gamma + cc * 32, g);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = (v[j] - mean) * rstd * g[j];
              add_vec32(st.ln_beta + cc * 32, v);
            }
            if (st.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (dst) store_act32(dst, row, cc, v);
            if (row_ok) {
              if (st.out_bf16) {
                uint32_t u[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) u[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
                __nv_bfloat16* o = static_cast<__nv_bfloat16*>(st.out_bf16) + (long long)m * st.ld_out_bf16 + cc * 32;
                st256b(o, u);
                st256b(o + 16, u + 8);
              }
              if (st.out_f32) {
                float* o = st.out_f32 + (long long)m * st.ld_out_f32 + cc * 32;
                if (st.out_f32_add) {         // e.g. the next residual = this output + the position feature (T:377-378)
                  float w[32];
#pragma unroll
                  for (int j = 0; j < 32; ++j) w[j] = v[j];
                  add_row32(st.out_f32_add + (long long)m * st.ld_out_f32_add + cc * 32, w);
#pragma unroll
                  for (int i = 0; i < 4; ++i) st256f(o + 8 * i, w + 8 * i);
                } else {
#pragma unroll
                  for (int i = 0; i < 4; ++i) st256f(o + 8 * i, v + 8 * i);
                }
              }
            }
            if (st.keep_col >= 0) {
              if (st.fold_bias) add_vec32(st.fold_bias + cc * 32, v);
              uint32_t q[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) q[j] = __float_as_uint(v[j]);
              tmem_st32_nowait(trow + (uint32_t)st.keep_col + cc * 32, q);
            }
          }
        }
      }
      if (st.keep_col >= 0) tmem_wait_st();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

inline bool al32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }
inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace tc

extern "C" int tc_linear_chain(const tc_chain_args* a, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(a != nullptr, TC_ERR_NULL, "tc_linear_chain: args is NULL");
  TC_REQUIRE(a->num_stages >= 1 && a->num_stages <= TC_CHAIN_MAX_STAGES, TC_ERR_SHAPE,
             "tc_linear_chain: num_stages must be 1..%d (got %d)", TC_CHAIN_MAX_STAGES, a->num_stages);
  TC_REQUIRE(a->M >= 0, TC_ERR_SHAPE, "tc_linear_chain: bad M");
  if (a->M == 0) return TC_OK;
  TC_REQUIRE(a->A != nullptr, TC_ERR_NULL, "tc_linear_chain: A is NULL");
  TC_REQUIRE(a->K >= CK && a->K <= 256 && a->K % CK == 0, TC_ERR_SHAPE, "tc_linear_chain: K must be 64, 128, 192 or 256 (got %d)", a->K);
  TC_REQUIRE(al16(a->A) && a->lda >= a->K && (a->lda * 2) % 16 == 0, TC_ERR_ALIGN, "tc_linear_chain: A rows must be 16-byte aligned");
  const int S = a->num_stages;
  ChainParams P;
  P.M = a->M; P.K0 = a->K; P.nstages = S;
  TC_REQUIRE(a->stage[0].K == a->K, TC_ERR_SHAPE, "tc_linear_chain: stage 0 K (%d) differs from A's K (%d)", a->stage[0].K, a->K);
  TC_REQUIRE(a->stage[S - 1].epi != TC_CHAIN_NONE, TC_ERR_SHAPE, "tc_linear_chain: the last stage needs an epilogue");
  for (int s = 0; s < S; ++s) {
    const tc_chain_stage& st = a->stage[s];
    TC_REQUIRE(st.W != nullptr, TC_ERR_NULL, "tc_linear_chain: stage %d: W is NULL", s);
    TC_REQUIRE(st.K >= CK && st.K <= 256 && st.K % CK == 0, TC_ERR_SHAPE, "tc_linear_chain: stage %d: K = %d unsupported", s, st.K);
    TC_REQUIRE(st.N >= 1 && st.N <= 256, TC_ERR_SHAPE, "tc_linear_chain: stage %d: N = %d unsupported", s, st.N);
    TC_REQUIRE(al16(st.W) && st.ldw >= st.K && (st.ldw * 2) % 16 == 0, TC_ERR_ALIGN, "tc_linear_chain: stage %d: W rows must be 16-byte aligned", s);
    TC_REQUIRE(st.a_buf == 0 || st.a_buf == 1, TC_ERR_SHAPE, "tc_linear_chain: stage %d: a_buf must be 0 or 1", s);
    const int n_mma = (st.N + 15) & ~15;
    TC_REQUIRE(st.acc_col >= 0 && st.acc_col % 32 == 0 && st.acc_col + ((st.N + 31) & ~31) <= 512, TC_ERR_SHAPE,
               "tc_linear_chain: stage %d: accumulator columns [%d, %d) out of range", s, st.acc_col, st.acc_col + n_mma);
    TC_REQUIRE(st.epi >= TC_CHAIN_NONE && st.epi <= TC_CHAIN_OUT, TC_ERR_SHAPE, "tc_linear_chain: stage %d: unknown epilogue", s);
    if (st.epi == TC_CHAIN_OUT) {
      TC_REQUIRE(st.N <= 32, TC_ERR_SHAPE, "tc_linear_chain: stage %d: TC_CHAIN_OUT needs N <= 32", s);
      TC_REQUIRE(st.out_f32 != nullptr && st.ld_out_f32 >= st.N, TC_ERR_NULL, "tc_linear_chain: stage %d: out_f32 missing", s);
      TC_REQUIRE(!st.row_bias || st.row_bias_period > 0, TC_ERR_SHAPE, "tc_linear_chain: stage %d: row_bias needs a period", s);
      TC_REQUIRE(st.tail >= TC_CHAIN_TAIL_NONE && st.tail <= TC_CHAIN_TAIL_ANCHOR_ADD, TC_ERR_SHAPE, "tc_linear_chain: stage %d: unknown tail", s);
      if (st.tail != TC_CHAIN_TAIL_NONE) {
        TC_REQUIRE(st.tail_in != nullptr && st.N >= 5, TC_ERR_NULL, "tc_linear_chain: stage %d: tail needs tail_in and N >= 5", s);
        TC_REQUIRE(st.tail != TC_CHAIN_TAIL_REF_UPDATE || (st.tail_out != nullptr && st.ld_tail_in >= 3), TC_ERR_NULL,
                   "tc_linear_chain: stage %d: ref update needs tail_out", s);
        TC_REQUIRE(st.tail != TC_CHAIN_TAIL_ANCHOR_ADD ||
                       (st.tail_xy_col >= 0 && st.tail_z_col >= 0 && st.ld_tail_in > st.tail_xy_col + 1 && st.ld_tail_in > st.tail_z_col),
                   TC_ERR_SHAPE, "tc_linear_chain: stage %d: anchor columns out of range", s);
      }
    } else {
      TC_REQUIRE(st.tail == TC_CHAIN_TAIL_NONE, TC_ERR_SHAPE, "tc_linear_chain: stage %d: a tail needs TC_CHAIN_OUT", s);
      TC_REQUIRE(st.N % 32 == 0, TC_ERR_SHAPE, "tc_linear_chain: stage %d: N must be a multiple of 32 (got %d)", s, st.N);
      if (st.epi != TC_CHAIN_NONE) {
        TC_REQUIRE(st.dst_buf >= -1 && st.dst_buf <= 1, TC_ERR_SHAPE, "tc_linear_chain: stage %d: dst_buf must be -1, 0 or 1", s);
        TC_REQUIRE(st.keep_col < 0 || (st.keep_col % 32 == 0 && st.keep_col + st.N <= 512), TC_ERR_SHAPE,
                   "tc_linear_chain: stage %d: keep_col out of range", s);
        TC_REQUIRE(st.epi != TC_CHAIN_LN || (st.ln_gamma && st.ln_beta), TC_ERR_NULL, "tc_linear_chain: stage %d: LayerNorm parameters missing", s);
        TC_REQUIRE(!st.out_f32 || (al32(st.out_f32) && st.ld_out_f32 % 8 == 0 && st.ld_out_f32 >= st.N), TC_ERR_ALIGN,
                   "tc_linear_chain: stage %d: out_f32 must be 32-byte aligned with a pitch multiple of 8", s);
        TC_REQUIRE(!st.out_f32_add || (st.out_f32 && al32(st.out_f32_add) && st.ld_out_f32_add % 8 == 0), TC_ERR_ALIGN,
                   "tc_linear_chain: stage %d: out_f32_add needs out_f32 and 32-byte aligned rows", s);
        TC_REQUIRE(!st.out_bf16 || (al32(st.out_bf16) && st.ld_out_bf16 % 16 == 0 && st.ld_out_bf16 >= st.N), TC_ERR_ALIGN,
                   "tc_linear_chain: stage %d: out_bf16 must be 32-byte aligned with a pitch multiple of 16", s);
        TC_REQUIRE(!st.bias || al16(st.bias), TC_ERR_ALIGN, "tc_linear_chain: stage %d: bias must be 16-byte aligned", s);
        TC_REQUIRE(!st.fold_bias || al16(st.fold_bias), TC_ERR_ALIGN, "tc_linear_chain: stage %d: fold_bias must be 16-byte aligned", s);
        TC_REQUIRE(!st.ln_gamma || (al16(st.ln_gamma) && al16(st.ln_beta)), TC_ERR_ALIGN, "tc_linear_chain: stage %d: LayerNorm parameters must be 16-byte aligned", s);
        TC_REQUIRE(st.dst_buf >= 0 || st.keep_col >= 0 || st.out_f32 || st.out_bf16, TC_ERR_NULL, "tc_linear_chain: stage %d: epilogue has no destination", s);
      }
    }
    if (st.init) {
      TC_REQUIRE(st.N % 32 == 0, TC_ERR_SHAPE, "tc_linear_chain: stage %d: init needs N %% 32 == 0", s);
      TC_REQUIRE(st.accumulate, TC_ERR_SHAPE, "tc_linear_chain: stage %d: init without accumulate would be overwritten", s);
      TC_REQUIRE(!st.init_bias || al16(st.init_bias), TC_ERR_ALIGN, "tc_linear_chain: stage %d: init_bias must be 16-byte aligned", s);
      TC_REQUIRE(!st.residual || (al32(st.residual) && st.ld_residual % 8 == 0), TC_ERR_ALIGN, "tc_linear_chain: stage %d: residual alignment", s);
      TC_REQUIRE(!st.residual2 || (al32(st.residual2) && st.ld_residual2 % 8 == 0), TC_ERR_ALIGN, "tc_linear_chain: stage %d: residual2 alignment", s);
    }
    if (s > 0) TC_REQUIRE(st.K <= 256, TC_ERR_SHAPE, "tc_linear_chain: stage %d: K too large", s);
    if (!tensor_map_bf16_2d(st.W, st.ldw, st.N, st.K, n_mma, &P.map_w[s])) return TC_ERR_SHAPE;
    P.st[s] = st;
  }
  if (!tensor_map_bf16_2d(a->A, a->lda, a->M, a->K, CM, &P.map_a)) return TC_ERR_SHAPE;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) { set_error("tc_linear_chain: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  cudaError_t e = launch(chain_kernel, dim3((unsigned)((a->M + CM - 1) / CM)), dim3(kThreads), kSmemBytes, as_stream(stream), 1u, P);
  if (e != cudaSuccess) { set_error("tc_linear_chain: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return (int)e; }
  count_launch();
  return check_launch("tc_linear_chain");
}
