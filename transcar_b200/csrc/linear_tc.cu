// K3 tensor-core path: Y = epilogue(A[M,K] * W[N,K]^T) with bf16 operands, fp32 accumulation in TMEM.
//
//   * operands: TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) into a 4-stage shared-memory ring,
//     full/empty mbarriers; both operands are K-major (nn.Linear keeps W as [N,K]).
//   * math: tcgen05.mma.cta_group::1.kind::f16, UMMA 128 x BN x 16 (BN = 64/128/256), issued by one thread;
//     accumulator = 128 lanes x BN fp32 columns of tensor memory.
//   * epilogue: 4 warps, one accumulator row per thread (tcgen05.ld 32x32b): bias / per-query row bias /
//     row gate / two residuals / LayerNorm over the full row (BN = N = 256: statistics are thread-local, the
//     pre-norm row is parked back in TMEM with tcgen05.st between the two passes) / ReLU / post-add, then fp32
//     and/or bf16 stores.  No intermediate ever goes to HBM.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue.
// One 128 x BN output tile per CTA (the GEMMs of this path are small: M = B*900 rows, N <= 1536, K <= 512;
// DESIGN.md discusses why they are latency- rather than throughput-bound).
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "tc_common.cuh"
#include "tc_sm100.cuh"

namespace tc {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int kStages = 4;
constexpr int kThreads = 192;

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
  int vec;        // 1: N % 32 == 0 and every row-wise operand is 16-byte aligned -> 128-bit epilogue accesses
};

// ---- epilogue math for one 32-column chunk of one row ------------------------------------------------
__device__ __forceinline__ void load32(const float* p, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(p + 4 * i);
    v[4 * i + 0] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}

template <int BN, bool kVec>
__global__ void __launch_bounds__(kThreads, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const EpiParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A 16 KB][B BN*128 B] then barriers
  constexpr uint32_t kABytes = BM * BK * 2;
  constexpr uint32_t kBBytes = BN * BK * 2;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  // pointer arithmetic (not an integer round-trip) so the compiler keeps the shared address space: LDS/STS, not LD/ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kStages), accbar = smem_u32(bars + 2 * kStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int num_kb = p.K / BK;

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

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        mbar_expect_tx(full0 + 8 * s, kStageBytes);
        const uint32_t a_dst = smem_base + s * kStageBytes;
        tma_load_2d(a_dst, &map_a, kb * BK, m0, full0 + 8 * s);
        tma_load_2d(a_dst + kABytes, &map_w, kb * BK, n0, full0 + 8 * s);
      }
    }
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
        const uint32_t a_addr = smem_base + s * kStageBytes;
        const uint64_t da = make_desc_sw128(a_addr), db = make_desc_sw128(a_addr + kABytes);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)          // +32 bytes (2 x 16 B) per UMMA_K step inside the 128 B row
          umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
        umma_commit(empty0 + 8 * s);               // frees the smem slot once these MMAs retire
      }
      umma_commit(accbar);                         // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====
    // Math is one accumulator row per thread (tcgen05.ld 32x32b); every global read/write of a row-major
    // operand goes through a per-warp 32x32 shared-memory tile so that each warp instruction touches whole
    // 128-byte row segments (the pipeline stages are free once the accumulator barrier has fired).
    const int quad = warp & 3;
    const int row0 = m0 + quad * 32;            // first row of this warp
    const int m = row0 + lane;
    const bool row_ok = m < p.M;
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    mbar_wait(accbar, 0);
    tc_fence_after();
    float (*tile)[33] = reinterpret_cast<float (*)[33]>(smem + (warp - 2) * 4352);
    constexpr bool vec = kVec;

    // coalesced [32 rows x 32 cols] fp32 tile -> this thread's row.  row r of the warp lives at base + rowoff(r).
    auto stage_in = [&](const float* base, long long ld, int period, int n, float (&t)[32]) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + (lane >> 3), c = (lane & 7) * 4;
        const int gm = row0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gm < p.M) {
          const float* src = base + (long long)(period > 0 ? gm % period : gm) * ld + n + c;
          if (vec) {
            v = *reinterpret_cast<const float4*>(src);
          } else {
            if (n + c + 0 < p.N) v.x = src[0];
            if (n + c + 1 < p.N) v.y = src[1];
            if (n + c + 2 < p.N) v.z = src[2];
            if (n + c + 3 < p.N) v.w = src[3];
          }
        }
        tile[r][c + 0] = v.x; tile[r][c + 1] = v.y; tile[r][c + 2] = v.z; tile[r][c + 3] = v.w;
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = tile[lane][j];
    };
    // per-column vector (bias / gamma / beta): same addresses for the whole warp -> broadcast loads
    auto load_vec = [&](const float* base, int n, float (&t)[32]) {
      if (vec) {
        load32(base + n, t);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = (n + j < p.N) ? base[n + j] : 0.f;
      }
    };
    auto stage_out = [&](int n, const float (&v)[32]) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; ++j) tile[lane][j] = v[j];
      __syncwarp();
      if (p.out_f32) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + (lane >> 3), c = (lane & 7) * 4;
          const int gm = row0 + r;
          if (gm < p.M) {
            float* dst = p.out_f32 + (long long)gm * p.ld_out_f32 + n + c;
            if (vec) {
              *reinterpret_cast<float4*>(dst) = make_float4(tile[r][c], tile[r][c + 1], tile[r][c + 2], tile[r][c + 3]);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (n + c + k < p.N) dst[k] = tile[r][c + k];
            }
          }
        }
      }
      if (p.out_bf16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 8 + (lane >> 2), c = (lane & 3) * 8;
          const int gm = row0 + r;
          if (gm < p.M) {
            __nv_bfloat16* dst = p.out_bf16 + (long long)gm * p.ld_out_bf16 + n + c;
            if (vec) {
              uint4 u;
              u.x = pack_bf16(tile[r][c + 0], tile[r][c + 1]); u.y = pack_bf16(tile[r][c + 2], tile[r][c + 3]);
              u.z = pack_bf16(tile[r][c + 4], tile[r][c + 5]); u.w = pack_bf16(tile[r][c + 6], tile[r][c + 7]);
              *reinterpret_cast<uint4*>(dst) = u;
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (n + c + k < p.N) dst[k] = __float2bfloat16_rn(tile[r][c + k]);
            }
          }
        }
      }
    };

    const bool gate = (p.row_gate && row_ok) ? (p.row_gate[m] != 0) : true;
    const bool do_ln = p.ln_gamma != nullptr;
    float s1 = 0.f, s2 = 0.f;

    // pass 1: accumulator -> pre-activation row value (bias, row bias, gate, residuals)
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(trow + c * 32, r);
      const int n = n0 + c * 32;
      float v[32], t[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (p.bias) { load_vec(p.bias, n, t);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += t[j]; }
      if (p.row_bias) { stage_in(p.row_bias, p.ld_row_bias, p.row_bias_period, n, t);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += t[j]; }
      if (!gate) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f; }
      if (p.residual) { stage_in(p.residual, p.ld_residual, 0, n, t);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += t[j]; }
      if (p.residual2) { stage_in(p.residual2, p.ld_residual2, 0, n, t);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += t[j]; }
      if (do_ln) {
#pragma unroll
        for (int j = 0; j < 32; ++j) { s1 += v[j]; s2 = fmaf(v[j], v[j], s2); r[j] = __float_as_uint(v[j]); }
        tmem_st32(trow + c * 32, r);          // park the pre-norm row in TMEM for pass 2
      } else {
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f); }
        if (p.post_add) { stage_in(p.post_add, p.ld_post_add, 0, n, t);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += t[j]; }
        stage_out(n, v);
      }
    }
    if (do_ln) {
      // pass 2: normalise (biased variance, like nn.LayerNorm), gain/bias, ReLU, post-add, store
      const float inv_n = 1.0f / (float)p.N;
      const float mean = s1 * inv_n;
      const float var = fmaxf(s2 * inv_n - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.ln_eps);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(trow + c * 32, r);
        const int n = n0 + c * 32;
        float v[32], g[32], b[32];
        load_vec(p.ln_gamma, n, g);
        load_vec(p.ln_beta, n, b);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = fmaf((__uint_as_float(r[j]) - mean) * rstd, g[j], b[j]);
          if (p.relu) v[j] = fmaxf(v[j], 0.f);
        }
        if (p.post_add) { stage_in(p.post_add, p.ld_post_add, 0, n, g);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += g[j]; }
        stage_out(n, v);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
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
  const void* ptr; long long ld; int rows, cols, box_rows;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && ld == o.ld && rows == o.rows && cols == o.cols && box_rows == o.box_rows;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h = h * 1000003u ^ (size_t)k.ld; h = h * 1000003u ^ (size_t)k.rows; h = h * 1000003u ^ (size_t)k.cols;
    return h * 1000003u ^ (size_t)k.box_rows;
  }
};

// bf16 [rows, cols] row-major with row stride ld (elements); box = [BK cols, box_rows rows], 128B swizzle,
// out-of-bounds rows read as zero.  Descriptors are pure functions of the key, so they are cached.
bool get_map(const void* ptr, long long ld, int rows, int cols, int box_rows, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, ld, rows, cols, box_rows};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("tc_linear: cuTensorMapEncodeTiled entry point not available"); return false; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("tc_linear: cuTensorMapEncodeTiled failed (%d)", (int)r); return false; }
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return true;
}

template <int BN, bool kVec>
int launch_bn(const tc_linear_args* a, const EpiParams& ep, cudaStream_t s) {
  CUtensorMap ma, mw;
  if (!get_map(a->A, a->lda, a->M, a->K, BM, &ma)) return TC_ERR_SHAPE;
  if (!get_map(a->W, a->ldw, a->N, a->K, BN, &mw)) return TC_ERR_SHAPE;
  constexpr size_t smem = (size_t)kStages * (BM * BK * 2 + BN * BK * 2) + 1024 /*align*/ + 256 /*barriers*/;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(linear_tc_kernel<BN, kVec>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("tc_linear: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  dim3 grid((a->M + BM - 1) / BM, (a->N + BN - 1) / BN);
  linear_tc_kernel<BN, kVec><<<grid, kThreads, smem, s>>>(ma, mw, ep);
  count_launch();
  return check_launch("tc_linear(tcgen05)");
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

static bool epilogue_vectorizable(const tc_linear_args* a) {
  if (a->N % 32 != 0) return false;
  if (a->bias && !al16(a->bias)) return false;
  if (a->row_bias && (!al16(a->row_bias) || a->ld_row_bias % 4 != 0)) return false;
  if (a->residual && (!al16(a->residual) || a->ld_residual % 4 != 0)) return false;
  if (a->residual2 && (!al16(a->residual2) || a->ld_residual2 % 4 != 0)) return false;
  if (a->post_add && (!al16(a->post_add) || a->ld_post_add % 4 != 0)) return false;
  if (a->ln_gamma && (!al16(a->ln_gamma) || !al16(a->ln_beta))) return false;
  if (a->out_f32 && (!al16(a->out_f32) || a->ld_out_f32 % 4 != 0)) return false;
  if (a->out_bf16 && (!al16(a->out_bf16) || a->ld_out_bf16 % 8 != 0)) return false;
  return true;
}

bool linear_tc_supported(const tc_linear_args* a) {
  static const bool disabled = getenv("TC_DISABLE_TC_LINEAR") != nullptr;        // debugging / A-B measurements
  if (disabled) return false;
  if (a->a_dtype != TC_BF16 || a->w_dtype != TC_BF16) return false;
  if (a->K % BK != 0 || a->K < BK) return false;
  const int N = a->N;
  if (!(N <= 32 || N == 64 || N == 128 || (N % 256 == 0))) return false;
  if (a->ln_gamma && N != 256) return false;
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
  const bool vec = epilogue_vectorizable(a);
  ep.vec = vec ? 1 : 0;
  if (a->N <= 32) return vec ? launch_bn<32, true>(a, ep, s) : launch_bn<32, false>(a, ep, s);
  if (a->N == 64) return vec ? launch_bn<64, true>(a, ep, s) : launch_bn<64, false>(a, ep, s);
  if (a->N == 128) return vec ? launch_bn<128, true>(a, ep, s) : launch_bn<128, false>(a, ep, s);
  return vec ? launch_bn<256, true>(a, ep, s) : launch_bn<256, false>(a, ep, s);
}

}  // namespace tc
