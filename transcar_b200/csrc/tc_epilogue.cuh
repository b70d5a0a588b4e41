// Epilogue helpers shared by the tensor-core Linear (linear_tc.cu) and the fused FFN (ffn_tc.cu): row-per-thread global
// access, 16-bit packing, cluster / distributed-shared-memory primitives of the LayerNorm statistics exchange, and the
// host-side tensor-map cache.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>

#include "tc_common.cuh"
#include "tc_sm100.cuh"

namespace tc {

// kinds of tensor map: the bf16 operand tiles, and the two output tiles of the TMA-store epilogue
enum MapKind { kMapOperand = 0, kMapOutF32 = 1, kMapOut16 = 2 };
// [rows, cols] row-major with row stride ld (elements), out-of-bounds rows / columns read as zero and are not written.
//   kMapOperand  bf16, box = [64 cols, box_rows rows], 128B swizzle (one swizzle row per K block row)
//   kMapOutF32   fp32, box = [32 cols, box_rows] = 128-byte rows, 128B swizzle      } the per-warp staging tiles of the
//   kMapOut16    16-bit, box = [32 cols, box_rows] = 64-byte rows, 64B swizzle      } epilogue (bf16 and fp16 alike)
// Descriptors are pure functions of the key, so they are cached (linear_tc.cu).
bool get_map(const void* ptr, long long ld, int rows, int cols, int box_rows, CUtensorMap* out, MapKind kind = kMapOperand);

// ---- tc_debug_trace (tools/linear_trace.py, tools/ffn_trace.py): per-CTA phase marks, 16 uint64 per CTA -------------------
// host: the next `ctas` records of the trace buffer, or null when tracing is off / the buffer is full (linear_tc.cu)
unsigned long long* trace_take(long long ctas);
static __device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
static __device__ __forceinline__ unsigned smid() {
  unsigned r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}
// compiled in only with -DTC_TRACE_BUILD (TC_TRACE_BUILD=1 python -m transcar_b200.build --force): the shipped kernels carry no marks
#ifdef TC_TRACE_BUILD
#define TC_TRACE(slot) do { if (trc) trc[slot] = (unsigned long long)clock64(); } while (0)
#else
#define TC_TRACE(slot) do { } while (0)
#endif

// ---- row-per-thread global access: 32 consecutive floats of one row ---------------------------------
static __device__ __forceinline__ void ld256(const float* p, float* v) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
static __device__ __forceinline__ void st256(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
static __device__ __forceinline__ void st256u(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// v[j] += row[j], j < 32 (ncols valid columns); 256-bit loads for every complete group of 8 columns
static __device__ __forceinline__ void add_row32(const float* row, int ncols, bool vec, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (vec && 8 * i + 8 <= ncols) {
      float t[8];
      ld256(row + 8 * i, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[8 * i + j] += t[j];
    } else {
#pragma unroll
      for (int j = 8 * i; j < 8 * i + 8; ++j)
        if (j < ncols) v[j] += row[j];
    }
  }
}

static __device__ __forceinline__ uint32_t pack_f16_sat(float lo, float hi) {       // saturating: +-65504 instead of inf
  lo = fminf(fmaxf(lo, -65504.f), 65504.f);
  hi = fminf(fmaxf(hi, -65504.f), 65504.f);
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

static __device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
static __device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
static __device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
// (no memory ordering: for warps that publish nothing but mbarrier inits, which fence.mbarrier_init.release.cluster covers)
static __device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
static __device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// LayerNorm statistics exchange between the CTAs of a cluster: every CTA PUSHES its per-row (sum, sum of squares) into
// slot [own rank][row] of each peer's shared memory (one relaxed 64-bit store: the value is its own flag, nothing else is
// published with it) and polls its own slots, which start as a NaN sentinel.  No cluster barrier sits on the critical path (the first version spent 41 % of its warp samples in
// barrier.cluster arrive / wait - ncu, profiles/r01_ncu_linear.txt): the only one (sentinels written before any peer
// may push) is split, arrive right after the init, wait just before the first push, ~5 us later.
// a NaN pair with a payload no arithmetic produces (sums of NaN inputs are the canonical 0x7fffffff): never a real value
constexpr unsigned long long kStatSentinel = 0xffc0dead'ffc0deadull;
static __device__ __forceinline__ void st_dsmem_u64(uint32_t local_addr, uint32_t rank, unsigned long long v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  asm volatile("st.relaxed.cluster.shared::cluster.b64 [%0], %1;" ::"r"(remote), "l"(v) : "memory");
}
static __device__ __forceinline__ unsigned long long ld_smem_u64(uint32_t addr) {
  unsigned long long v;
  asm volatile("ld.relaxed.cluster.shared::cta.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}

// ---- cluster pair helpers of the fused kernels (ffn_tc.cu) ------------------------------------------------
// arrive on the same-offset mbarrier of CTA `rank` of the cluster
static __device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// bulk copy of `bytes` (multiple of 16) from this CTA's shared memory into the same-layout shared memory of CTA `rank`;
// completion is counted (complete_tx) on that CTA's mbarrier at offset `local_bar`
static __device__ __forceinline__ void bulk_copy_to_cluster(uint32_t dst_local, uint32_t src, uint32_t bytes, uint32_t local_bar, uint32_t rank) {
  uint32_t rdst, rbar;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdst) : "r"(dst_local), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(local_bar), "r"(rank));
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(rdst), "r"(src), "r"(bytes), "r"(rbar) : "memory");
}
static __device__ __forceinline__ void lds128(uint32_t addr, float* v) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr) : "memory");
}

}  // namespace tc
