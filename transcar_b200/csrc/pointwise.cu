// Small fused pointwise stages of the fusion decoder and the device-side NMS-free decode (N1).
#include "tc_common.cuh"

namespace tc {
namespace {

// T:195-203: new_ref = sigmoid(code[{0,1,4}] + inverse_sigmoid(ref))
__global__ void ref_update_kernel(const float* __restrict__ code, long long ld_code, const float* __restrict__ ref,
                                  float* __restrict__ out, int M) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * 3) return;
  const int m = i / 3, a = i % 3;
  const int col = a == 2 ? 4 : a;
  out[i] = sigmoid_f32(__fadd_rn(code[(long long)m * ld_code + col], logit_f32(ref[i])));
}

struct AnchorParams {
  float* code; long long ld_code; const float* anchor; long long ld_anchor;
  int xy_col, z_col, from_norm, M; float pc[6];
};
__global__ void box_anchor_add_kernel(const AnchorParams p) {
  pdl_trigger();
  pdl_wait();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= p.M) return;
  const float* an = p.anchor + (long long)m * p.ld_anchor;
  float ax = an[p.xy_col], ay = an[p.xy_col + 1];
  const float az = an[p.z_col];
  if (p.from_norm) {   // H:596-597; z is added un-scaled (quirk Q3, H:598 is an empty slice)
    ax = __fadd_rn(__fmul_rn(ax, p.pc[3] - p.pc[0]), p.pc[0]);
    ay = __fadd_rn(__fmul_rn(ay, p.pc[4] - p.pc[1]), p.pc[1]);
  }
  float* c = p.code + (long long)m * p.ld_code;
  c[0] = __fadd_rn(c[0], ax);
  c[1] = __fadd_rn(c[1], ay);
  c[4] = __fadd_rn(c[4], az);
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, long long ld_src, __nv_bfloat16* __restrict__ dst,
                                 long long ld_dst, int rows, int cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  dst[(long long)r * ld_dst + c] = __float2bfloat16_rn(src[(long long)r * ld_src + c]);
}

__global__ void cast_split_kernel(const float* __restrict__ src, long long ld_src, __nv_bfloat16* __restrict__ dst,
                                  long long ld_dst, int rows, int cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  const float v = src[(long long)r * ld_src + c];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  dst[(long long)r * ld_dst + c] = hi;
  dst[(long long)r * ld_dst + cols + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// ---- N1 decode: one block per sample, bitonic sort of (score, index) keys in shared memory ----------
// key = orderable(score) << 32 | ~index  ->  descending sort gives scores high-to-low, ties by low index.
__device__ __forceinline__ uint32_t orderable(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct DecodeParams {
  const float* cls; const float* code; int Q, classes, max_num, npad; float rng[6];
  float* boxes; float* scores; int* labels; uint8_t* keep; float* records;
};

__global__ void __launch_bounds__(1024) decode_kernel(const DecodeParams p) {
  extern __shared__ unsigned long long keys[];
  const int b = blockIdx.x;
  const int n = p.Q * p.classes;
  const float* cls = p.cls + (long long)b * n;
  for (int i = threadIdx.x; i < p.npad; i += blockDim.x) {
    unsigned long long k = 0ull;          // padding sorts last
    if (i < n) k = ((unsigned long long)orderable(sigmoid_f32(cls[i])) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
    keys[i] = k;
  }
  __syncthreads();
  for (int size = 2; size <= p.npad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < p.npad / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = keys[lo], c = keys[hi];
        if ((a < c) == desc) { keys[lo] = c; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int r = threadIdx.x; r < p.max_num; r += blockDim.x) {
    const long long o = (long long)b * p.max_num + r;
    float bx[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float score = 0.f;
    int label = 0;
    bool in = false;
    if (r < n) {
      const unsigned long long k = keys[r];
      const uint32_t idx = 0xffffffffu - (uint32_t)(k & 0xffffffffull);
      const int qi = idx / p.classes;
      const float* c = p.code + ((long long)b * p.Q + qi) * 10;
      // U:26-52 denormalize: (cx,cy,w,l,cz,h,sin,cos,vx,vy) -> (cx,cy,cz,w,l,h,rot,vx,vy)
      const float cx = c[0], cy = c[1], cz = c[4];
      bx[0] = cx; bx[1] = cy; bx[2] = cz;
      bx[3] = expf(c[2]); bx[4] = expf(c[3]); bx[5] = expf(c[5]);
      bx[6] = atan2f(c[6], c[7]);
      bx[7] = c[8]; bx[8] = c[9];
      score = sigmoid_f32(cls[idx]);
      label = (int)(idx % p.classes);
      in = cx >= p.rng[0] && cy >= p.rng[1] && cz >= p.rng[2] && cx <= p.rng[3] && cy <= p.rng[4] && cz <= p.rng[5];
    }
    if (p.boxes) {
      for (int j = 0; j < 9; ++j) p.boxes[o * 9 + j] = bx[j];
      p.scores[o] = score; p.labels[o] = label; p.keep[o] = in ? 1 : 0;
    }
    if (p.records) {
      float* rec = p.records + o * 12;
      for (int j = 0; j < 9; ++j) rec[j] = bx[j];
      rec[9] = score; rec[10] = (float)label; rec[11] = in ? 1.f : 0.f;
    }
  }
}

static int next_pow2(int n) { int p = 2; while (p < n) p <<= 1; return p; }

}  // namespace
}  // namespace tc

extern "C" int tc_ref_update(const float* code, int64_t ld_code, const float* ref, float* new_ref, int32_t M,
                             tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(code && ref && new_ref, TC_ERR_NULL, "tc_ref_update: NULL pointer");
  TC_REQUIRE(M >= 0 && ld_code >= 5, TC_ERR_SHAPE, "tc_ref_update: bad shape");
  if (M == 0) return TC_OK;
  launch(ref_update_kernel, dim3((M * 3 + 255) / 256), dim3(256), 0, as_stream(stream), 1u, code, ld_code, ref, new_ref, M);
  count_launch();
  return check_launch("tc_ref_update");
}

extern "C" int tc_box_anchor_add(float* code, int64_t ld_code, const float* anchor, int64_t ld_anchor, int32_t xy_col,
                                 int32_t z_col, int32_t xy_from_normalised, const float* pc_range6, int32_t M,
                                 tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(code && anchor && pc_range6, TC_ERR_NULL, "tc_box_anchor_add: NULL pointer");
  TC_REQUIRE(M >= 0 && ld_code >= 5 && xy_col >= 0 && z_col >= 0 && ld_anchor > xy_col + 1 && ld_anchor > z_col,
             TC_ERR_SHAPE, "tc_box_anchor_add: bad shape");
  if (M == 0) return TC_OK;
  AnchorParams p{code, ld_code, anchor, ld_anchor, xy_col, z_col, xy_from_normalised, M, {}};
  for (int i = 0; i < 6; ++i) p.pc[i] = pc_range6[i];
  launch(box_anchor_add_kernel, dim3((M + 127) / 128), dim3(128), 0, as_stream(stream), 1u, p);
  count_launch();
  return check_launch("tc_box_anchor_add");
}

extern "C" int tc_cast_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int32_t rows, int32_t cols,
                            tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(src && dst, TC_ERR_NULL, "tc_cast_bf16: NULL pointer");
  TC_REQUIRE(rows >= 0 && cols >= 0 && ld_src >= cols && ld_dst >= cols, TC_ERR_SHAPE, "tc_cast_bf16: bad shape");
  const long long n = (long long)rows * cols;
  if (n == 0) return TC_OK;
  cast_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
      src, ld_src, static_cast<__nv_bfloat16*>(dst), ld_dst, rows, cols);
  count_launch();
  return check_launch("tc_cast_bf16");
}

extern "C" int tc_cast_split(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int32_t rows, int32_t cols,
                             tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(src && dst, TC_ERR_NULL, "tc_cast_split: NULL pointer");
  TC_REQUIRE(rows >= 0 && cols >= 0 && ld_src >= cols && ld_dst >= 2 * (int64_t)cols, TC_ERR_SHAPE, "tc_cast_split: bad shape");
  const long long n = (long long)rows * cols;
  if (n == 0) return TC_OK;
  cast_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
      src, ld_src, static_cast<__nv_bfloat16*>(dst), ld_dst, rows, cols);
  count_launch();
  return check_launch("tc_cast_split");
}

extern "C" int64_t tc_decode_workspace_bytes(int32_t B, int32_t Q, int32_t classes) {
  (void)B; (void)Q; (void)classes;
  return 0;   // the sort runs in shared memory
}

extern "C" int tc_decode(const tc_decode_args* a, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(a != nullptr, TC_ERR_NULL, "tc_decode: args is NULL");
  TC_REQUIRE(a->cls && a->code, TC_ERR_NULL, "tc_decode: NULL input pointer");
  const bool quad = a->boxes && a->scores && a->labels && a->keep;
  TC_REQUIRE(quad || (!a->boxes && !a->scores && !a->labels && !a->keep), TC_ERR_NULL,
             "tc_decode: boxes / scores / labels / keep go together");
  TC_REQUIRE(quad || a->records, TC_ERR_NULL, "tc_decode: no output pointer");
  TC_REQUIRE(a->B >= 0 && a->Q > 0 && a->classes > 0 && a->max_num > 0, TC_ERR_SHAPE, "tc_decode: bad shape");
  const int n = a->Q * a->classes;
  const int npad = next_pow2(n);
  TC_REQUIRE((size_t)npad * 8 <= 200 * 1024, TC_ERR_SHAPE, "tc_decode: Q*classes = %d too large for the in-smem sort", n);
  if (a->B == 0) return TC_OK;
  DecodeParams p{a->cls, a->code, a->Q, a->classes, a->max_num, npad, {}, a->boxes, a->scores, a->labels, a->keep, a->records};
  for (int i = 0; i < 6; ++i) p.rng[i] = a->post_center_range[i];
  const size_t smem = (size_t)npad * 8;
  cudaError_t e = cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("tc_decode: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  decode_kernel<<<a->B, 1024, smem, as_stream(stream)>>>(p);
  count_launch();
  return check_launch("tc_decode");
}
