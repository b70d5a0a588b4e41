// K1: fused DETR3D camera sampling (projection + validity mask + 4-level bilinear gather + masked
// sigmoid-weighted reduction).  See include/transcar_b200.h (tc_sample_fwd) for the contract and the
// reference lines it replaces (detr3d_transformer.py:367-373, 381-422).
//
// HBM-bound gather: algorithmic bytes per valid (query, camera) pair = 16 texels * C * sizeof(feat).
// A random 8 KB gather is latency-bound unless ~100 KB per SM is in flight; a register-resident design (16 warps x
// 16 x 128-bit loads, round-1 v1) spends half of its time with nothing in flight and reached 51 % of the HBM
// roofline.  Here the texels are staged through shared memory with cp.async (LDGSTS: no registers held while in
// flight) and the issue of a pair is decoupled from its consumption by a per-warp ring:
//
//   * one warp per (sample, query), persistent warps: a CTA owns a contiguous range of queries and hands them to its
//     warps through a shared-memory counter (valid cameras per query vary 0..3, a static assignment leaves a tail);
//     the next query's reference point / logits / camera matrices are prefetched into registers one query ahead;
//   * lane c < N projects the query through camera c's lidar2img (fp32, reference op order, bit-exact mask),
//     a ballot yields the valid-camera set (~18 % of the pairs are valid);
//   * for every valid pair and every 512-byte channel chunk, lanes 0..15 each own ONE texel (level = lane/4,
//     corner = lane%4): they compute its index and bilinear*sigmoid weight and publish both in shared memory;
//     then the whole warp issues 16 cp.async (one texel each, 16 bytes per lane = 512 coalesced bytes) into the
//     warp's ring slot (8 KB) and commits them as one group.  A lane later reads back exactly the 16-byte pieces
//     it copied itself, so the texel data needs no cross-lane synchronisation at all;
//   * a warp keeps kSlots pairs in flight; it consumes the oldest slot (cp.async.wait_group, 16 conflict-free
//     128-bit shared-memory reads per lane, fp32 FMA into the lane's 8 (bf16) or 4 (fp32) channels) only when the
//     ring is full, and stores a query's channels once its last camera has been consumed.
// (A cp.async.bulk per texel was measured first: UBLKCP takes uniform registers, so 16 per-lane copies turn into a
// serialised ELECT/R2UR loop that cost a third of the kernel.)
// 12 warps x 2 slots x 8 KB = 192 KB of shared memory per CTA, one CTA per SM.  Measured and rejected (round 1): a global atomic query counter (7200 same-address atomics
// cost more than the tail they remove) and the 16-texel weighted sum as mma.sync m16n8k16 with 3-way bf16-split weights
// (bit-accurate, but the accumulator fragment leaves 1 lane in 4 with useful data: the per-query epilogue cost more
// than the 192 FMA/unpack instructions it replaced; 13-17 us vs 10.5 us).
#include <cstdlib>

#include "tc_common.cuh"

namespace tc {
namespace {

constexpr int kTexels = 16;                     // 4 levels x 4 corners
constexpr int kChunkBytes = 512;                // bytes of one texel handled per pass: 256 bf16 or 128 fp32 channels
constexpr int kSlotBytes = kTexels * kChunkBytes;
// shared memory of a CTA with `warps` warps of `slots` ring slots each: ring + per-slot weights, texel indices, meta
constexpr int sample_smem(int warps, int slots) { return warps * slots * (kSlotBytes + kTexels * 4 + kTexels * 4 + 16) + 16; }

struct SampleParams {
  const void* feat[TC_MAX_LEVELS];
  int H[TC_MAX_LEVELS];
  int W[TC_MAX_LEVELS];
  int B, N, Q, C;
  const float* ref;
  const float* lidar2img;
  const float* logits;
  float pc[6];
  float inv_w, inv_h;      // 1.0f / img_w, 1.0f / img_h
  float inv_q;             // 1.0f / Q
  void* out;
  uint8_t* mask;
  int all_cams;            // TC_SAMPLE_ALL_CAMS
  int weights_given;       // TC_SAMPLE_WEIGHTS_GIVEN
};

// ATen grid_sampler_unnormalize, align_corners=False: ((g + 1) * size - 1) / 2.
__device__ __forceinline__ float unnormalize(float g, int size) {
  return (__fadd_rn(g, 1.0f) * static_cast<float>(size) - 1.0f) * 0.5f;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most `n` of this thread's committed groups are still pending
__device__ __forceinline__ void cp_async_wait(unsigned n) {
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
  }
}

// query row of task t: floor(t / Q) without an integer division (t, Q < 2^30; inv_q = 1.0f / Q)
__device__ __forceinline__ int row_of(int t, int Q, float inv_q) {
  int b = (int)((float)t * inv_q);
  b += ((b + 1) * Q <= t) ? 1 : 0;
  b -= (b * Q > t) ? 1 : 0;
  return b;
}

// Per-query inputs, fetched one query ahead of their use.
struct QueryIn {
  float rx, ry, rz, logit;
  float4 m0, m1, m2;          // rows 0..2 of this lane's camera matrix (lane < N)
};

__device__ __forceinline__ void fetch_query(const SampleParams& p, int task, int lane, QueryIn& in) {
  const int b = row_of(task, p.Q, p.inv_q);
  const float* r = p.ref + (size_t)task * 3;
  in.rx = __ldg(r); in.ry = __ldg(r + 1); in.rz = __ldg(r + 2);            // same address in every lane: broadcast
  in.logit = (lane < p.N * 4) ? __ldg(p.logits + (size_t)task * (p.N * 4) + lane) : 0.f;
  if (lane < p.N) {
    const float4* m = reinterpret_cast<const float4*>(p.lidar2img + ((size_t)b * p.N + lane) * 16);
    in.m0 = __ldg(m); in.m1 = __ldg(m + 1); in.m2 = __ldg(m + 2);
  }
}

// kOut: 0 = fp32 [.., C], 1 = bf16 [.., C], 2 = split bf16 [.., 2C] (hi | lo: the A operand of a bf16x3 Linear)
template <bool kBf16In, int kOut, int kWarps, int kSlots>
__global__ void __launch_bounds__(kWarps * 32, kWarps <= 6 ? 2 : 1) sample_kernel(const SampleParams p) {
  constexpr bool kBf16Out = kOut != 0;
  constexpr int kOutRow = kOut == 2 ? 2 : 1;           // output row pitch in units of C
  constexpr int kPer = kBf16In ? 8 : 4;                 // channels per lane per chunk (16 bytes)
  constexpr int kChunkCh = kBf16In ? 256 : 128;         // channels per 512-byte chunk
  constexpr int kRingBytes = kWarps * kSlots * kSlotBytes;
  constexpr int kWeightBytes = kWarps * kSlots * kTexels * 4;          // bilinear * sigmoid weight per texel
  constexpr int kIndexBytes = kWarps * kSlots * kTexels * 4;           // texel index per texel
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* ring = smem + warp * kSlots * kSlotBytes + lane * 16;          // this lane's 16-byte column of the ring
  float* wts = reinterpret_cast<float*>(smem + kRingBytes) + warp * kSlots * kTexels;
  unsigned* tidx = reinterpret_cast<unsigned*>(smem + kRingBytes + kWeightBytes) + warp * kSlots * kTexels;
  int* meta = reinterpret_cast<int*>(smem + kRingBytes + kWeightBytes + kIndexBytes) + warp * kSlots * 4;
  unsigned* next_ctr = reinterpret_cast<unsigned*>(smem + kRingBytes + kWeightBytes + kIndexBytes + kWarps * kSlots * 16);
  const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);

  const int total = p.B * p.Q;
  const int esz = kBf16In ? 2 : 4;
  const int chunks = p.C / kChunkCh;
  unsigned issued = 0, consumed = 0;                    // ring counters (warp-uniform)
  float acc[kPer];
#pragma unroll
  for (int i = 0; i < kPer; ++i) acc[i] = 0.f;

  // ---- consume the oldest slot: wait for its 8 KB, weighted sum into acc, store on the pair that ends a (query, chunk)
  auto consume = [&]() {
    const unsigned slot = consumed % kSlots;
    cp_async_wait(issued - consumed - 1);                                  // the oldest group has landed
    const int4 mt = *reinterpret_cast<const int4*>(meta + slot * 4);        // row, channel offset, flags
    if (mt.z & 1) {
#pragma unroll
      for (int i = 0; i < kPer; ++i) acc[i] = 0.f;
    }
    const uint8_t* src = ring + slot * kSlotBytes;
    const float4* w4 = reinterpret_cast<const float4*>(wts + slot * kTexels);
#pragma unroll
    for (int t4 = 0; t4 < kTexels / 4; ++t4) {
      const float4 w = w4[t4];
      const float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + (t4 * 4 + j) * kChunkBytes);
        if (kBf16In) {
          acc[0] = fmaf(bf16_lo(v.x), ws[j], acc[0]); acc[1] = fmaf(bf16_hi(v.x), ws[j], acc[1]);
          acc[2] = fmaf(bf16_lo(v.y), ws[j], acc[2]); acc[3] = fmaf(bf16_hi(v.y), ws[j], acc[3]);
          acc[4 % kPer] = fmaf(bf16_lo(v.z), ws[j], acc[4 % kPer]); acc[5 % kPer] = fmaf(bf16_hi(v.z), ws[j], acc[5 % kPer]);
          acc[6 % kPer] = fmaf(bf16_lo(v.w), ws[j], acc[6 % kPer]); acc[7 % kPer] = fmaf(bf16_hi(v.w), ws[j], acc[7 % kPer]);
        } else {
          acc[0] = fmaf(__uint_as_float(v.x), ws[j], acc[0]); acc[1] = fmaf(__uint_as_float(v.y), ws[j], acc[1]);
          acc[2] = fmaf(__uint_as_float(v.z), ws[j], acc[2]); acc[3] = fmaf(__uint_as_float(v.w), ws[j], acc[3]);
        }
      }
    }
    if (mt.z & 2) {
      const size_t o = (size_t)mt.x * p.C * kOutRow + mt.y + lane * kPer;
      if (kBf16Out) {
        __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.out) + o;
        if (kBf16In) {
          uint4 u;
          u.x = pack_bf16(acc[0], acc[1]); u.y = pack_bf16(acc[2], acc[3]);
          u.z = pack_bf16(acc[4 % kPer], acc[5 % kPer]); u.w = pack_bf16(acc[6 % kPer], acc[7 % kPer]);
          *reinterpret_cast<uint4*>(dst) = u;
          if (kOut == 2) {
            uint4 l;
            l.x = pack_bf16(acc[0] - bf16_lo(u.x), acc[1] - bf16_hi(u.x)); l.y = pack_bf16(acc[2] - bf16_lo(u.y), acc[3] - bf16_hi(u.y));
            l.z = pack_bf16(acc[4 % kPer] - bf16_lo(u.z), acc[5 % kPer] - bf16_hi(u.z));
            l.w = pack_bf16(acc[6 % kPer] - bf16_lo(u.w), acc[7 % kPer] - bf16_hi(u.w));
            *reinterpret_cast<uint4*>(dst + p.C) = l;
          }
        } else {
          uint2 u;
          u.x = pack_bf16(acc[0], acc[1]); u.y = pack_bf16(acc[2], acc[3]);
          *reinterpret_cast<uint2*>(dst) = u;
          if (kOut == 2) {
            uint2 l;
            l.x = pack_bf16(acc[0] - bf16_lo(u.x), acc[1] - bf16_hi(u.x)); l.y = pack_bf16(acc[2] - bf16_lo(u.y), acc[3] - bf16_hi(u.y));
            *reinterpret_cast<uint2*>(dst + p.C) = l;
          }
        }
      } else {
        float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + o);
        dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        if (kBf16In) dst[1] = make_float4(acc[4 % kPer], acc[5 % kPer], acc[6 % kPer], acc[7 % kPer]);
      }
    }
    ++consumed;
    __syncwarp();                      // every lane has read the slot's weights / meta before they are overwritten
  };

  // Work distribution: a CTA owns a contiguous range of queries; its warps take the first kWarps statically and the
  // rest from a shared-memory counter (the number of valid cameras per query varies 0..3, so a static assignment
  // leaves some warps with twice the average work).  A warp knows its next query one iteration ahead (`next`), so
  // that query's inputs are prefetched.
  const int range_begin = (int)(((long long)total * blockIdx.x) / gridDim.x);
  const int range_end = (int)(((long long)total * (blockIdx.x + 1)) / gridDim.x);
  if (threadIdx.x == 0) *next_ctr = kWarps;
  __syncthreads();
  auto grab = [&]() -> int {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(next_ctr, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    return min(range_end, range_begin + (int)t);
  };
  int task = min(range_end, range_begin + warp);
  QueryIn cur, nxt;
  pdl_trigger();
  pdl_wait();            // reference points and logits come from the previous kernels
  if (task < range_end) fetch_query(p, task, lane, cur);
  nxt = cur;
  int next = task < range_end ? grab() : range_end;

  for (; task < range_end; task = next, next = (next < range_end ? grab() : range_end), cur = nxt) {
    const int b = row_of(task, p.Q, p.inv_q);

    // -- projection: lane c handles camera c (T:389-409), fp32, reference op order ------------------
    const float px = __fadd_rn(__fmul_rn(cur.rx, p.pc[3] - p.pc[0]), p.pc[0]);
    const float py = __fadd_rn(__fmul_rn(cur.ry, p.pc[4] - p.pc[1]), p.pc[1]);
    const float pz = __fadd_rn(__fmul_rn(cur.rz, p.pc[5] - p.pc[2]), p.pc[2]);
    float gx = 0.f, gy = 0.f;
    bool valid = false;
    if (lane < p.N) {
      // 4x4 . (px,py,pz,1) in the evaluation order of the batched K=4 SGEMM that torch.matmul dispatches to on
      // sm_100 (found by tools/probe_matmul4.py: bit-identical on 21600/21600 outputs):
      //   (m0*px (+) m1*py)  +  (m2*pz (+) m3*1)   with (+) = FMA
      const float4 m0 = cur.m0, m1 = cur.m1, m2 = cur.m2;
      float cx = __fadd_rn(__fmaf_rn(m0.y, py, __fmul_rn(m0.x, px)), __fmaf_rn(m0.w, 1.0f, __fmul_rn(m0.z, pz)));
      float cy = __fadd_rn(__fmaf_rn(m1.y, py, __fmul_rn(m1.x, px)), __fmaf_rn(m1.w, 1.0f, __fmul_rn(m1.z, pz)));
      float cz = __fadd_rn(__fmaf_rn(m2.y, py, __fmul_rn(m2.x, px)), __fmaf_rn(m2.w, 1.0f, __fmul_rn(m2.z, pz)));
      const float eps = 1e-5f;
      valid = cz > eps;
      const float zc = fmaxf(cz, eps);
      // `x /= python_scalar` on a CUDA tensor is x * (1/scalar) in ATen (BinaryDivTrueKernel: cpu-scalar fast path),
      // not a true division; tensor / tensor (the perspective divide) is IEEE division.
      float u = __fmul_rn(__fdiv_rn(cx, zc), p.inv_w);
      float v = __fmul_rn(__fdiv_rn(cy, zc), p.inv_h);
      gx = __fmul_rn(__fadd_rn(u, -0.5f), 2.0f);
      gy = __fmul_rn(__fadd_rn(v, -0.5f), 2.0f);
      valid = valid && (gx > -1.0f) && (gx < 1.0f) && (gy > -1.0f) && (gy < 1.0f);
      if (p.mask) p.mask[(size_t)task * p.N + lane] = valid ? 1 : 0;
    }
    const unsigned vset = __ballot_sync(0xffffffffu, valid || (p.all_cams && lane < p.N));

    // -- sigmoid(attention logits): lane i < N*4 holds weight i = cam*4 + level ----------------------
    const float wgt = (lane < p.N * 4) ? (p.weights_given ? cur.logit : sigmoid_f32(cur.logit)) : 0.f;

    // -- next query's inputs: issued now, consumed in the next iteration ------------------------------
    if (next < range_end) fetch_query(p, next, lane, nxt);

    if (vset == 0) {                                    // no camera sees the point: the masked sum is exactly 0
      for (int c = lane * kPer; c < p.C * kOutRow; c += 32 * kPer) {
        const size_t o = (size_t)task * p.C * kOutRow + c;
        if (kBf16Out) {
          if (kPer == 8) *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + o) = make_uint4(0, 0, 0, 0);
          else *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out) + o) = make_uint2(0, 0);
        } else {
          float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + o);
          dst[0] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (kPer == 8) dst[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      continue;
    }

    const int level = (lane >> 2) & 3, corner = lane & 3;
    const int H = p.H[level], W = p.W[level];
    const int last_cam = 31 - __clz(vset);
    for (int chunk = 0; chunk < chunks; ++chunk) {
      unsigned todo = vset;
      bool first = true;
      while (todo) {
        const int cam = __ffs(todo) - 1;
        todo &= todo - 1;
        if (issued - consumed == kSlots) consume();
        // ---- issue: lanes 0..15 each own one texel of this (query, camera)
        const float x = __shfl_sync(0xffffffffu, gx, cam);
        const float y = __shfl_sync(0xffffffffu, gy, cam);
        const float lw = __shfl_sync(0xffffffffu, wgt, cam * 4 + level);
        const float ix = unnormalize(x, W), iy = unnormalize(y, H);
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const int xi = (int)fx0 + (corner & 1), yi = (int)fy0 + (corner >> 1);      // nw, ne, sw, se
        // ATen corner weights: (ix_se - ix) / (ix - ix_nw) and likewise in y
        const float wx = (corner & 1) ? ix - fx0 : (fx0 + 1.0f) - ix;
        const float wy = (corner >> 1) ? iy - fy0 : (fy0 + 1.0f) - iy;
        const bool inside = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H);
        // zeros padding: out-of-range corners get weight 0 and a clamped (always legal) address
        const float w = inside ? wx * wy * lw : 0.f;
        const int cxi = min(max(xi, 0), W - 1), cyi = min(max(yi, 0), H - 1);
        const unsigned texel = (unsigned)((((size_t)b * p.N + cam) * H + cyi) * W + cxi);
        const unsigned slot = issued % kSlots;
        if (lane < kTexels) {
          wts[slot * kTexels + lane] = w;
          tidx[slot * kTexels + lane] = texel;
        }
        if (lane == 0)
          *reinterpret_cast<int4*>(meta + slot * 4) =
              make_int4(task, chunk * kChunkCh, (first ? 1 : 0) | (cam == last_cam ? 2 : 0), 0);
        __syncwarp();
        // ---- all lanes: 16 x (512 coalesced bytes global -> this lane's ring column)
        const size_t row_bytes = (size_t)p.C * esz;
        const size_t lane_off = (size_t)chunk * kChunkBytes + lane * 16;
        const uint32_t dst = ring_u32 + slot * kSlotBytes;
#pragma unroll
        for (int t4 = 0; t4 < kTexels / 4; ++t4) {
          const uint4 ti = *reinterpret_cast<const uint4*>(tidx + slot * kTexels + t4 * 4);
          const uint8_t* base = static_cast<const uint8_t*>(p.feat[t4]) + lane_off;
          cp_async16(dst + (t4 * 4 + 0) * kChunkBytes, base + ti.x * row_bytes);
          cp_async16(dst + (t4 * 4 + 1) * kChunkBytes, base + ti.y * row_bytes);
          cp_async16(dst + (t4 * 4 + 2) * kChunkBytes, base + ti.z * row_bytes);
          cp_async16(dst + (t4 * 4 + 3) * kChunkBytes, base + ti.w * row_bytes);
        }
        cp_async_commit();
        ++issued;
        first = false;
      }
    }
  }
  while (consumed != issued) consume();
}

// ---- NCHW fp32 -> NHWC (fp32|bf16) tiled transpose ------------------------------------------------
template <bool kBf16Out>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ src, void* __restrict__ dst,
                                                            int C, int HW) {
  __shared__ float tile[32][33];
  const size_t plane = blockIdx.z;
  const int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    int c = c0 + i, hw = hw0 + tx;
    tile[i][tx] = (c < C && hw < HW) ? src[(plane * C + c) * (size_t)HW + hw] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int hw = hw0 + i, c = c0 + tx;
    if (hw < HW && c < C) {
      float v = tile[tx][i];
      size_t o = (plane * HW + hw) * (size_t)C + c;
      if (kBf16Out) static_cast<__nv_bfloat16*>(dst)[o] = __float2bfloat16_rn(v);
      else static_cast<float*>(dst)[o] = v;
    }
  }
}

template <bool kBf16In, int kOut, int kWarps, int kSlots>
cudaError_t launch_sample(const SampleParams& p, long long total, int sm_count, cudaStream_t s) {
  constexpr int smem = sample_smem(kWarps, kSlots);
  static bool configured = false;
  auto kernel = sample_kernel<kBf16In, kOut, kWarps, kSlots>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const long long ctas = (total + kWarps - 1) / kWarps;
  const long long resident = (long long)sm_count * (kWarps <= 6 ? 2 : 1);       // 6-warp CTAs (96 KB): two per SM
  return launch(kernel, dim3((unsigned)(ctas < resident ? ctas : resident)), dim3(kWarps * 32), (size_t)smem, s, 1u, p);
}

// Tuning hook (tools/k1_bench.py): TC_SAMPLE_VARIANT picks the warps x ring-slots shape of the bf16 kernel.
int sample_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TC_SAMPLE_VARIANT");
    v = e ? atoi(e) : 0;
  }
  return v;
}

}  // namespace
}  // namespace tc

extern "C" int tc_sample_fwd(const tc_sample_args* a, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(a != nullptr, TC_ERR_NULL, "tc_sample_fwd: args is NULL");
  if (a->B == 0 || a->Q == 0) return TC_OK;         // empty batch / query set: nothing to do
  TC_REQUIRE(a->ref && a->lidar2img && a->attn_logits && a->out, TC_ERR_NULL, "tc_sample_fwd: NULL tensor pointer");
  TC_REQUIRE(a->num_levels == 4, TC_ERR_SHAPE, "tc_sample_fwd: num_levels must be 4 (got %d)", a->num_levels);
  TC_REQUIRE(a->N >= 1 && a->N <= TC_MAX_CAMS && a->N * a->num_levels <= 32, TC_ERR_SHAPE,
             "tc_sample_fwd: num cams %d unsupported", a->N);
  TC_REQUIRE(a->C > 0 && a->C % 256 == 0, TC_ERR_SHAPE, "tc_sample_fwd: C must be a multiple of 256 (got %d)", a->C);
  // B * Q < 2^22: the kernel recovers the batch index of a row with a float reciprocal (+-1 fix-up), exact in that range
  TC_REQUIRE(a->B >= 0 && a->Q >= 0 && (long long)a->B * a->Q < (1ll << 22), TC_ERR_SHAPE, "tc_sample_fwd: bad B/Q (B * Q must be < 2^22)");
  TC_REQUIRE(aligned16(a->lidar2img), TC_ERR_ALIGN, "tc_sample_fwd: lidar2img must be 16-byte aligned");
  TC_REQUIRE(a->feat_dtype == TC_F32 || a->feat_dtype == TC_BF16, TC_ERR_DTYPE, "tc_sample_fwd: bad feat dtype");
  TC_REQUIRE(a->out_dtype == TC_F32 || a->out_dtype == TC_BF16 || a->out_dtype == TC_BF16X2, TC_ERR_DTYPE,
             "tc_sample_fwd: bad out dtype");
  TC_REQUIRE(aligned16(a->out), TC_ERR_ALIGN, "tc_sample_fwd: out must be 16-byte aligned");
  if (a->B == 0 || a->Q == 0) return TC_OK;
  SampleParams p;
  for (int l = 0; l < 4; ++l) {
    TC_REQUIRE(a->feat[l] != nullptr, TC_ERR_NULL, "tc_sample_fwd: feat[%d] is NULL", l);
    TC_REQUIRE(aligned16(a->feat[l]), TC_ERR_ALIGN, "tc_sample_fwd: feat[%d] must be 16-byte aligned", l);
    TC_REQUIRE(a->H[l] > 0 && a->W[l] > 0, TC_ERR_SHAPE, "tc_sample_fwd: level %d has empty extent", l);
    p.feat[l] = a->feat[l]; p.H[l] = a->H[l]; p.W[l] = a->W[l];
  }
  p.B = a->B; p.N = a->N; p.Q = a->Q; p.C = a->C;
  p.ref = a->ref; p.lidar2img = a->lidar2img; p.logits = a->attn_logits;
  for (int i = 0; i < 6; ++i) p.pc[i] = a->pc_range[i];
  p.inv_w = 1.0f / a->img_w; p.inv_h = 1.0f / a->img_h; p.inv_q = 1.0f / (float)a->Q;
  p.out = a->out; p.mask = a->mask;
  p.all_cams = (a->flags & TC_SAMPLE_ALL_CAMS) ? 1 : 0;
  p.weights_given = (a->flags & TC_SAMPLE_WEIGHTS_GIVEN) ? 1 : 0;
  // persistent warps, one CTA per SM; each CTA hands its range of queries out to its warps dynamically
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    sm_count = n;
  }
  cudaStream_t s = as_stream(stream);
  const bool bi = a->feat_dtype == TC_BF16, bo = a->out_dtype == TC_BF16, so = a->out_dtype == TC_BF16X2;
  const long long total = (long long)a->B * a->Q;
  cudaError_t e = cudaSuccess;
  // bf16 in / bf16 out (the engine's path): 24 warps x 1 slot is faster back to back (10.5 us vs 10.9 us for 12 x 2,
  // tools/k1_bench.py: the launch ramp hides behind the previous launch) but slower inside the step (bracket 19.8 us vs
  // 18.4 us, step +9 us: 768-thread CTAs start later), so 12 x 2 stays the default.
  if (bi && bo) {
    if (sample_variant() == 1) e = launch_sample<true, 1, 24, 1>(p, total, sm_count, s);
    else if (sample_variant() == 2) e = launch_sample<true, 1, 6, 2>(p, total, sm_count, s);
    else e = launch_sample<true, 1, 12, 2>(p, total, sm_count, s);
    if (false) e = launch_sample<true, 1, 6, 2>(p, total, sm_count, s);
  } else if (bi && so) {
    if (sample_variant() == 2) e = launch_sample<true, 2, 6, 2>(p, total, sm_count, s);
    else e = launch_sample<true, 2, 12, 2>(p, total, sm_count, s);
  }
  else if (bi) e = launch_sample<true, 0, 12, 2>(p, total, sm_count, s);
  else if (so) e = launch_sample<false, 2, 12, 2>(p, total, sm_count, s);
  else if (bo) e = launch_sample<false, 1, 12, 2>(p, total, sm_count, s);
  else e = launch_sample<false, 0, 12, 2>(p, total, sm_count, s);
  if (e != cudaSuccess) { set_error("tc_sample_fwd: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return (int)e; }
  count_launch();
  return check_launch("tc_sample_fwd");
}

extern "C" int tc_nchw_to_nhwc(const float* src, void* dst, int32_t dst_dtype, int32_t planes, int32_t C,
                               int32_t HW, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(src && dst, TC_ERR_NULL, "tc_nchw_to_nhwc: NULL pointer");
  TC_REQUIRE(planes >= 0 && C > 0 && HW > 0 && planes <= 65535, TC_ERR_SHAPE, "tc_nchw_to_nhwc: bad shape");
  TC_REQUIRE(dst_dtype == TC_F32 || dst_dtype == TC_BF16, TC_ERR_DTYPE, "tc_nchw_to_nhwc: bad dtype");
  if (planes == 0) return TC_OK;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, planes);
  if (dst_dtype == TC_BF16) nchw_to_nhwc_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(src, dst, C, HW);
  else nchw_to_nhwc_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(src, dst, C, HW);
  count_launch();
  return check_launch("tc_nchw_to_nhwc");
}
