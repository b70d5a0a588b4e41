// Shared device/host helpers for the transcar_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/transcar_b200.h"

namespace tc {

// ---- host-side error plumbing (capi.cu owns the storage) -------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);     // returns cudaGetLastError() (0 = ok) and records the message

#define TC_REQUIRE(cond, code, ...)            \
  do {                                         \
    if (!(cond)) {                             \
      ::tc::set_error(__VA_ARGS__);            \
      return (code);                           \
    }                                          \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline cudaStream_t as_stream(tc_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------
// The step is ~130 small kernels in one stream / CUDA graph; each kernel lets its successor start early
// (pdl_trigger at entry) and waits for its predecessor only where it first touches activations (pdl_wait), so
// launch latency and per-CTA prologues (barrier init, TMEM allocation, parameter staging) overlap the previous
// kernel's tail.  Rule for every kernel: nothing produced by an earlier kernel is read, and nothing at all is
// written to global memory, before pdl_wait().  TC_DISABLE_PDL=1 turns the launch attribute off (A/B runs).
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                          unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// torch.sigmoid in fp32: 1 / (1 + exp(-x)) with the full-precision expf.
__device__ __forceinline__ float sigmoid_f32(float x) { return 1.0f / (1.0f + expf(-x)); }

// inverse_sigmoid (T:17-32): clamp to [0,1], eps 1e-5 on numerator and denominator, log of the ratio.
__device__ __forceinline__ float logit_f32(float x) {
  x = fminf(fmaxf(x, 0.0f), 1.0f);
  float a = fmaxf(x, 1e-5f);
  float b = fmaxf(__fsub_rn(1.0f, x), 1e-5f);
  return logf(__fdiv_rn(a, b));
}

// 128-bit read-only global load that does not allocate in L1 (streaming gathers).
__device__ __forceinline__ uint4 ldg_nc_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ---- radar distance mask: one (query circle set, radar point) test --------------------------------
// torch.cdist(p=2) on [900,2]x[1500,2] takes the matmul route (_euclidean_dist):
//   x_ = [-2*x0, -2*x1, |x|^2, 1]   y_ = [y0, y1, 1, |y|^2]   d = sqrt(clamp_min(x_ . y_, 0))
// The dot product is restated as an fp32 FMA chain in k order (what a K=4 SGEMM does); the norms are
// rounded squares added without contraction (pow(2) and sum are separate ATen kernels).
struct Circle { float m2x, m2y, nrm; };     // -2*cx, -2*cy, cx^2+cy^2

__device__ __forceinline__ Circle make_circle(float cx, float cy) {
  Circle c;
  c.m2x = __fmul_rn(cx, -2.0f);
  c.m2y = __fmul_rn(cy, -2.0f);
  c.nrm = __fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy));
  return c;
}
__device__ __forceinline__ float key_norm(float kx, float ky) {
  return __fadd_rn(__fmul_rn(kx, kx), __fmul_rn(ky, ky));
}
__device__ __forceinline__ float cdist_mm(const Circle& c, float kx, float ky, float knrm) {
  float acc = __fmul_rn(c.m2x, kx);
  acc = __fmaf_rn(c.m2y, ky, acc);
  acc = __fmaf_rn(c.nrm, 1.0f, acc);
  acc = __fmaf_rn(1.0f, knrm, acc);
  return __fsqrt_rn(acc < 0.0f ? 0.0f : acc);       // clamp_min(0) as torch.cdist does it: a NaN stays NaN (never < radius)
}
// geom = (cx, cy, fx, fy, rx, ry, radius, -)
__device__ __forceinline__ bool radar_allowed(const Circle& c, const Circle& f, const Circle& r, float radius,
                                              float kx, float ky, float knrm) {
  return (cdist_mm(c, kx, ky, knrm) < radius) | (cdist_mm(f, kx, ky, knrm) < radius) |
         (cdist_mm(r, kx, ky, knrm) < radius);
}

// ---- counter-based dropout (training variant) ---------------------------------------------------------------------
// Philox4x32-10 keyed by (seed, stream); element (row, col) of a logical 2-D tensor takes component col & 3 of the block with
// counter (col >> 2, row).  A mask is a pure function of (seed, stream, row, col): the backward kernels regenerate it
// instead of storing it, and tc_dropout on a tensor of ones materialises it for the parity tests.  Semantics of
// nn.Dropout / the attention-probability dropout of nn.MultiheadAttention (H:122-145, H:578): keep with probability 1 - p,
// scale kept values by 1 / (1 - p).
struct DropoutRng {
  float p, inv_keep;
  uint32_t seed_lo, seed_hi, stream_lo, stream_hi;
};
inline DropoutRng make_rng(float p, uint64_t seed, uint64_t stream) {
  DropoutRng r;
  r.p = p; r.inv_keep = p < 1.f ? 1.0f / (1.0f - p) : 0.f;
  r.seed_lo = (uint32_t)seed; r.seed_hi = (uint32_t)(seed >> 32);
  r.stream_lo = (uint32_t)stream; r.stream_hi = (uint32_t)(stream >> 32);
  return r;
}
__device__ __forceinline__ uint4 philox4(uint32_t c0, uint32_t c1, const DropoutRng& r) {
  uint32_t k0 = r.seed_lo, k1 = r.seed_hi;
  uint4 c = make_uint4(c0, c1, r.stream_lo, r.stream_hi);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ bool keep_from(uint32_t bits, float p) { return (float)(bits >> 8) * (1.0f / 16777216.0f) >= p; }
// multiplier of element (row, col): 0 or 1 / (1 - p)
__device__ __forceinline__ float dropout_scale(const DropoutRng& r, uint32_t row, uint32_t col) {
  const uint4 x = philox4(col >> 2, row, r);
  const uint32_t bits = (col & 2) ? ((col & 1) ? x.w : x.z) : ((col & 1) ? x.y : x.x);
  return keep_from(bits, r.p) ? r.inv_keep : 0.f;
}

// ---- per-query circle geometry of the radar mask (H:543-567 / H:615-635 / H:671-693) -----------------------------
// centre (cx, cy) in metres, box code columns 3 (log length), 6, 7 (heading terms) -> g[0..8) =
// (cx, cy, fx, fy, rx, ry, radius, thr); thr = smallest fp32 y with sqrt_rn(y) >= radius, so that
// sqrt(x) < radius  <=>  x < thr exactly (x >= 0).  ONE definition for tc_radar_geometry and the fused Linear tails.
__device__ __forceinline__ void radar_geom_row(float cx, float cy, float code3, float code6, float code7, float r_lo, float r_hi,
                                               float* g) {
  const float len = expf(code3);            // H:553
  const float s = -code6, c = -code7;       // H:554-555
  const float ox = __fmul_rn(__fmul_rn(len, 0.25f), s);     // object_length*0.25*object_rot_sin (left-assoc.)
  const float oy = __fmul_rn(__fmul_rn(len, 0.25f), c);
  g[0] = cx; g[1] = cy;
  g[2] = __fadd_rn(cx, ox); g[3] = __fadd_rn(cy, oy);
  g[4] = __fsub_rn(cx, ox); g[5] = __fsub_rn(cy, oy);
  const float radius = fminf(fmaxf(__fdiv_rn(len, 2.0f), r_lo), r_hi);
  g[6] = radius;
  float thr = __fmul_rn(radius, radius);
  for (int it = 0; it < 8 && thr > 0.f && __fsqrt_rn(thr) >= radius; ++it) thr = nextafterf(thr, 0.f);
  for (int it = 0; it < 16 && __fsqrt_rn(thr) < radius; ++it) thr = nextafterf(thr, INFINITY);
  g[7] = thr;
}

// Row-local tails of tc_linear (include/transcar_b200.h): y = the row's first 8 outputs (columns 0..7), updated in place
// for TC_TAIL_BOX.  Same operations, in the same order, as ref_update_kernel / box_anchor_add_kernel / radar_geometry_kernel.
struct TailParams {
  int kind;
  const float* in; long long ld_in;
  float* ref_out; float* geom_out;
  int xy_col, z_col, from_norm;
  float pc[6]; float r_lo, r_hi;
};
__device__ __forceinline__ void apply_tail(const TailParams& t, long long m, float* y) {
  const float* in = t.in + m * t.ld_in;
  if (t.kind == TC_TAIL_REF_UPDATE) {
    const float r0 = sigmoid_f32(__fadd_rn(y[0], logit_f32(in[0])));
    const float r1 = sigmoid_f32(__fadd_rn(y[1], logit_f32(in[1])));
    const float r2 = sigmoid_f32(__fadd_rn(y[4], logit_f32(in[2])));
    t.ref_out[m * 3 + 0] = r0; t.ref_out[m * 3 + 1] = r1; t.ref_out[m * 3 + 2] = r2;
    if (t.geom_out) {
      const float cx = __fadd_rn(__fmul_rn(r0, t.pc[3] - t.pc[0]), t.pc[0]);
      const float cy = __fadd_rn(__fmul_rn(r1, t.pc[4] - t.pc[1]), t.pc[1]);
      float g[8];
      radar_geom_row(cx, cy, y[3], y[6], y[7], t.r_lo, t.r_hi, g);
      float4* o = reinterpret_cast<float4*>(t.geom_out + m * 8);
      o[0] = make_float4(g[0], g[1], g[2], g[3]);
      o[1] = make_float4(g[4], g[5], g[6], g[7]);
    }
  } else if (t.kind == TC_TAIL_BOX) {
    float ax = in[t.xy_col], ay = in[t.xy_col + 1];
    const float az = in[t.z_col];
    if (t.from_norm) {   // H:596-597; z is added un-scaled (quirk Q3, H:598 is an empty slice)
      ax = __fadd_rn(__fmul_rn(ax, t.pc[3] - t.pc[0]), t.pc[0]);
      ay = __fadd_rn(__fmul_rn(ay, t.pc[4] - t.pc[1]), t.pc[1]);
    }
    y[0] = __fadd_rn(y[0], ax);
    y[1] = __fadd_rn(y[1], ay);
    y[4] = __fadd_rn(y[4], az);
    if (t.geom_out) {
      float g[8];
      radar_geom_row(y[0], y[1], y[3], y[6], y[7], t.r_lo, t.r_hi, g);
      float4* o = reinterpret_cast<float4*>(t.geom_out + m * 8);
      o[0] = make_float4(g[0], g[1], g[2], g[3]);
      o[1] = make_float4(g[4], g[5], g[6], g[7]);
    }
  }
}
inline TailParams make_tail(const tc_linear_args* a) {
  TailParams t;
  t.kind = a->tail; t.in = a->tail_in; t.ld_in = a->ld_tail_in; t.ref_out = a->tail_ref_out; t.geom_out = a->tail_geom_out;
  t.xy_col = a->tail_xy_col; t.z_col = a->tail_z_col; t.from_norm = a->tail_from_norm;
  for (int i = 0; i < 6; ++i) t.pc[i] = a->tail_pc_range[i];
  t.r_lo = a->tail_r_lo; t.r_hi = a->tail_r_hi;
  return t;
}

}  // namespace tc
