// K3w wide-tile variant of the tensor-core Linear for plain large-N products in the bf16x3 format: Y = A W^T + bias (ReLU
// optional), N a multiple of 256 - the stacked radar K / V projection of the three radar layers (M = 12 000, N = 1 536,
// K = 256: in_proj rows C.. of rf_multihead_attn{,2,3}, H:578 / H:646 / H:704).
//
// With 128 x 64 tiles that GEMM is 2 256 CTAs whose A tiles are re-read 24 times (433 MB through L2, 50-65 us), and because it
// runs on a side branch its queued CTAs stand in front of whatever the critical path launches next (~40 us per step, CUPTI
// timeline).  Here a CTA owns a 128 x 256 tile: UMMA 128 x 256 x 16 (three bf16 passes per K step), 2 x 96 KB ring, 564 CTAs,
// half the operand bytes.  Epilogue: eight warps, one accumulator row x 32 columns at a time per thread, staged in the idle
// ring as swizzled 32 x 32 tiles and written by TMA tile stores (fp32 and / or one 16-bit format: bf16 or fp16).  A per-query
// row bias is supported, so the self-attention in-projection (N = 768: q / k / v of mmcv's MultiheadAttention wrapper, the
// query_pos half cached as row bias) can take the same path (171 CTAs).
#include <cuda.h>

#include <cstdlib>

#include "tc_common.cuh"
#include "tc_epilogue.cuh"
#include "tc_sm100.cuh"

namespace tc {
namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int kThreads = 320;
constexpr uint32_t kATile = BM * BK * 2;                       // 16 KB
constexpr uint32_t kWTile = BN * BK * 2;                       // 32 KB
constexpr uint32_t kStageBytes = 2 * kATile + 2 * kWTile;      // A_hi | A_lo | W_hi | W_lo = 96 KB
constexpr int kStages = 2;
constexpr uint32_t kOffBars = kStages * kStageBytes;           // full[2], empty[2], acc, tmem slot
constexpr uint32_t kOffVec = kOffBars + 64;                    // bias (BN)
constexpr uint32_t kSmemUsed = kOffVec + BN * 4;
constexpr size_t kSmemBytes = kSmemUsed + 1024;

struct WideParams {
  int M, N, K;
  const float* bias;
  const float* row_bias; int row_bias_period; long long ld_row_bias;
  int relu;
  int w_static;
  int has_o32, out16;       // out16: 0 = none, TC_BF16 or TC_F16
};

__global__ void __launch_bounds__(kThreads, 1)
linear_wide_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_o32, const __grid_constant__ CUtensorMap map_o16, const WideParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* s_bias = reinterpret_cast<float*>(smem + kOffVec);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + 2), accbar = smem_u32(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int num_kb = p.K / BK;
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
  const int num_pre = p.w_static ? min(num_kb, kStages) : 0;     // static weights: requested before the dependency wait
  if (threadIdx.x == 0) {
    for (int kb = 0; kb < num_pre; ++kb) {
      const uint32_t w_dst = smem_base + kb * kStageBytes + 2 * kATile;
      mbar_expect_tx(full0 + 8 * kb, kStageBytes);
      tma_load_2d(w_dst, &map_w, kb * BK, n0, full0 + 8 * kb);
      tma_load_2d(w_dst + kWTile, &map_w, p.K + kb * BK, n0, full0 + 8 * kb);
    }
  }
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        const uint32_t st = smem_base + s * kStageBytes, full = full0 + 8 * s;
        const bool w_done = kb < num_pre;
        if (!w_done) {
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          mbar_expect_tx(full, kStageBytes);
        }
        tma_load_2d(st, &map_a, kb * BK, m0, full);
        tma_load_2d(st + kATile, &map_a, p.K + kb * BK, m0, full);
        if (!w_done) {
          tma_load_2d(st + 2 * kATile, &map_w, kb * BK, n0, full);
          tma_load_2d(st + 2 * kATile + kWTile, &map_w, p.K + kb * BK, n0, full);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(full0 + 8 * s, ph);
        tc_fence_after();
        const uint32_t st = smem_base + s * kStageBytes;
        const uint64_t dah = make_desc_sw128(st), dal = make_desc_sw128(st + kATile);
        const uint64_t dwh = make_desc_sw128(st + 2 * kATile), dwl = make_desc_sw128(st + 2 * kATile + kWTile);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_base, dah + 2 * k, dwh + 2 * k, idesc, (kb | k) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_base, dal + 2 * k, dwh + 2 * k, idesc, 1u);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_base, dah + 2 * k, dwl + 2 * k, idesc, 1u);
        umma_commit(empty0 + 8 * s);
      }
      umma_commit(accbar);
    }
    __syncwarp();
  } else {
    // ===== epilogue: TMEM lane quadrant = warp % 4, column half h = (warp - 2) / 4: four 32-column chunks per thread =====
    const int quad = warp & 3, h = (warp - 2) >> 2;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    {
      const int e = threadIdx.x - 64;                  // 0..255
      s_bias[e] = p.bias ? p.bias[n0 + e] : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    const int row = quad * 32 + lane, m = m0 + row;
    const bool row_ok = m < p.M;
    const float* rb = (p.row_bias && row_ok) ? p.row_bias + (long long)(m % p.row_bias_period) * p.ld_row_bias + n0 + 128 * h : nullptr;
    float side[32];                                    // per-query row bias of the next chunk, requested one chunk ahead
    auto fetch_side = [&](int c) {
      if (rb) {
#pragma unroll
        for (int i = 0; i < 4; ++i) ld256(rb + 32 * c + 8 * i, side + 8 * i);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) side[j] = 0.f;
      }
    };
    fetch_side(0);
    mbar_wait(accbar, 0);
    tc_fence_after();
    const int mrow = m0 + quad * 32;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const int col = 128 * h + 32 * c;                // first of this chunk's columns inside the tile
      uint32_t r[32];
      tmem_ld32(tlane + (uint32_t)col, r);
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = __uint_as_float(r[j]) + s_bias[col + j] + side[j];
        if (p.relu) v[j] = fmaxf(v[j], 0.f);
      }
      if (c + 1 < 4) fetch_side(c + 1);
      // two staging tiles (fp32 4 KB | 16-bit 2 KB) per warp, used alternately: the tile of chunk c - 2 has been read by now
      const uint32_t stg = smem_base + (uint32_t)((warp - 2) * 2 + (c & 1)) * 6144u;
      if (c >= 2 && lane == 0) tma_store_wait_read1();
      __syncwarp();
      if (p.has_o32) {
        const uint32_t rowa = stg + (uint32_t)lane * 128u;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          sts128(rowa + (uint32_t)((q ^ (lane & 7)) << 4), __float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]),
                 __float_as_uint(v[4 * q + 2]), __float_as_uint(v[4 * q + 3]));
      }
      if (p.out16) {
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
        for (int q = 0; q < 4; ++q) sts128(rowa + (uint32_t)((q ^ sw) << 4), u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && mrow < p.M) {
        if (p.has_o32) tma_store_2d(&map_o32, stg, n0 + col, mrow);
        if (p.out16) tma_store_2d(&map_o16, stg + 4096u, n0 + col, mrow);
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait_read();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

// plain large-N bf16x3 products (see the header of this file)
bool linear_wide_supported(const tc_linear_args* a) {
  static const bool disabled = getenv("TC_NO_WIDE_LINEAR") != nullptr;            // A/B measurements
  static const int min_tiles = getenv("TC_WIDE_MIN_TILES") ? atoi(getenv("TC_WIDE_MIN_TILES")) : 3;
  if (disabled) return false;
  if (a->a_dtype != TC_BF16X2 || a->w_dtype != TC_BF16X2) return false;
  if (a->N % BN != 0 || a->N < min_tiles * BN || a->K % BK != 0 || a->K < BK || a->M < 16 * BM) return false;
  if (a->row_gate || a->residual || a->residual2 || a->ln_gamma || a->post_add || a->tail) return false;
  if (a->row_bias && ((reinterpret_cast<uintptr_t>(a->row_bias) & 31u) || a->ld_row_bias % 8 != 0)) return false;
  const int o16 = a->out16_dtype == 0 ? TC_BF16 : a->out16_dtype;
  if (a->out_bf16 && (o16 == TC_BF16X2 || !al16(a->out_bf16) || (a->ld_out_bf16 * 2) % 16 != 0)) return false;
  if (a->out_f32 && (!al16(a->out_f32) || (a->ld_out_f32 * 4) % 16 != 0)) return false;
  return al16(a->A) && al16(a->W) && (a->lda * 2) % 16 == 0 && (a->ldw * 2) % 16 == 0;
}

int linear_wide_launch(const tc_linear_args* a, cudaStream_t s) {
  CUtensorMap ma, mw, mo, mo16;
  if (!get_map(a->A, a->lda, a->M, 2 * a->K, BM, &ma)) return TC_ERR_SHAPE;
  if (!get_map(a->W, a->ldw, a->N, 2 * a->K, BN, &mw)) return TC_ERR_SHAPE;
  mo = ma; mo16 = ma;
  if (a->out_f32 && !get_map(a->out_f32, a->ld_out_f32, a->M, a->N, 32, &mo, kMapOutF32)) return TC_ERR_SHAPE;
  if (a->out_bf16 && !get_map(a->out_bf16, a->ld_out_bf16, a->M, a->N, 32, &mo16, kMapOut16)) return TC_ERR_SHAPE;
  WideParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.bias = a->bias; p.relu = a->relu; p.w_static = a->w_static ? 1 : 0;
  p.row_bias = a->row_bias; p.row_bias_period = a->row_bias_period > 0 ? a->row_bias_period : 1; p.ld_row_bias = a->ld_row_bias;
  p.has_o32 = a->out_f32 ? 1 : 0;
  p.out16 = a->out_bf16 ? (a->out16_dtype == 0 ? TC_BF16 : a->out16_dtype) : 0;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(linear_wide_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) { set_error("tc_linear(wide): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  cudaError_t e = launch(linear_wide_tc_kernel, dim3((unsigned)(a->N / BN), (unsigned)((a->M + BM - 1) / BM)), dim3(kThreads), kSmemBytes, s,
                         1u, ma, mw, mo, mo16, p);
  if (e != cudaSuccess) { set_error("tc_linear(wide): %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return (int)e; }
  count_launch();
  return check_launch("tc_linear(wide)");
}

}  // namespace tc
