// C-ABI plumbing: error state, launch accounting, argument validation and path dispatch for
// tc_linear / tc_attention_fwd.  See include/transcar_b200.h.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "tc_common.cuh"

namespace tc {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return TC_OK;
}

bool pdl_enabled() {
  static const bool on = getenv("TC_DISABLE_PDL") == nullptr;
  return on;
}

int linear_simt_launch(const tc_linear_args* a, cudaStream_t s);
int attention_simt_launch(const tc_attention_args* a, cudaStream_t s);
bool linear_wide_supported(const tc_linear_args* a);
int linear_wide_launch(const tc_linear_args* a, cudaStream_t s);
bool linear_tc_supported(const tc_linear_args* a);
int linear_tc_launch(const tc_linear_args* a, cudaStream_t s);
bool attention_tc_supported(const tc_attention_args* a);
int attention_tc_launch(const tc_attention_args* a, cudaStream_t s);
bool attention_sparse_supported(const tc_attention_args* a);
int attention_sparse_launch(const tc_attention_args* a, cudaStream_t s);

}  // namespace tc

extern "C" int tc_abi_version(void) { return TC_ABI_VERSION; }
extern "C" const char* tc_last_error_string(void) { return tc::g_err; }
extern "C" uint64_t tc_launch_count(void) { return tc::g_launches.load(std::memory_order_relaxed); }

extern "C" int tc_check_device(void) {
  int dev = 0, major = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) {
    tc::set_error("tc_check_device: %s", cudaGetErrorString(e));
    return (int)e;
  }
  if (major != 10) {
    tc::set_error("tc_check_device: compute capability %d.x is not sm_100 (this library is B200-only)", major);
    return TC_ERR_DEVICE;
  }
  return TC_OK;
}

extern "C" int tc_linear(const tc_linear_args* a, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(a != nullptr, TC_ERR_NULL, "tc_linear: args is NULL");
  TC_REQUIRE(a->M >= 0 && a->N > 0 && a->K > 0, TC_ERR_SHAPE, "tc_linear: bad M/N/K %d/%d/%d", a->M, a->N, a->K);
  TC_REQUIRE(!a->ln_gamma || a->N <= 256, TC_ERR_SHAPE, "tc_linear: fused LayerNorm needs N <= 256 (got %d)", a->N);
  if (a->M == 0) return TC_OK;                      // empty row set: nothing to do (pointers may be NULL)
  TC_REQUIRE(a->A && a->W, TC_ERR_NULL, "tc_linear: A or W is NULL");
  TC_REQUIRE(a->out_f32 || a->out_bf16, TC_ERR_NULL, "tc_linear: no output pointer");
  TC_REQUIRE(a->M >= 0 && a->N > 0 && a->K > 0, TC_ERR_SHAPE, "tc_linear: bad M/N/K %d/%d/%d", a->M, a->N, a->K);
  const bool split_a = a->a_dtype == TC_BF16X2, split_w = a->w_dtype == TC_BF16X2;
  TC_REQUIRE((a->a_dtype == TC_F32 || a->a_dtype == TC_BF16 || split_a) && (a->w_dtype == TC_F32 || a->w_dtype == TC_BF16 || split_w),
             TC_ERR_DTYPE, "tc_linear: bad dtype");
  TC_REQUIRE(split_a == split_w, TC_ERR_DTYPE, "tc_linear: split bf16 (TC_BF16X2) operands come in pairs");
  TC_REQUIRE(a->lda >= (split_a ? 2 : 1) * (int64_t)a->K && a->ldw >= (split_w ? 2 : 1) * (int64_t)a->K, TC_ERR_SHAPE,
             "tc_linear: leading dimension smaller than K (2K for split operands)");
  TC_REQUIRE(a->out16_dtype == 0 || a->out16_dtype == TC_BF16 || a->out16_dtype == TC_BF16X2 || a->out16_dtype == TC_F16,
             TC_ERR_DTYPE, "tc_linear: bad out16_dtype %d", a->out16_dtype);
  TC_REQUIRE(!a->ln_gamma || a->ln_beta, TC_ERR_NULL, "tc_linear: ln_gamma without ln_beta");
  TC_REQUIRE(!a->ln_gamma || a->N <= 256, TC_ERR_SHAPE, "tc_linear: fused LayerNorm needs N <= 256 (got %d)", a->N);
  TC_REQUIRE(!a->row_bias || a->row_bias_period > 0, TC_ERR_SHAPE, "tc_linear: row_bias needs a positive period");
  TC_REQUIRE(!a->out_f32 || a->ld_out_f32 >= a->N, TC_ERR_SHAPE, "tc_linear: ld_out_f32 < N");
  TC_REQUIRE(!a->out_bf16 || a->ld_out_bf16 >= (a->out16_dtype == TC_BF16X2 ? 2 : 1) * (int64_t)a->N, TC_ERR_SHAPE,
             "tc_linear: ld_out_bf16 < N (2N for a split output)");
  if (a->tail != TC_TAIL_NONE) {
    TC_REQUIRE(a->tail == TC_TAIL_REF_UPDATE || a->tail == TC_TAIL_BOX, TC_ERR_SHAPE, "tc_linear: unknown tail %d", a->tail);
    TC_REQUIRE(a->N >= 8 && a->N <= 32 && !a->ln_gamma && !a->relu, TC_ERR_SHAPE,
               "tc_linear: a tail needs 8 <= N <= 32, no LayerNorm and no ReLU (got N = %d)", a->N);
    TC_REQUIRE(a->tail_in, TC_ERR_NULL, "tc_linear: tail_in is NULL");
    TC_REQUIRE(a->tail != TC_TAIL_REF_UPDATE || (a->tail_ref_out && a->ld_tail_in >= 3), TC_ERR_NULL,
               "tc_linear: TC_TAIL_REF_UPDATE needs tail_ref_out and ld_tail_in >= 3");
    TC_REQUIRE(a->tail != TC_TAIL_BOX || (a->tail_xy_col >= 0 && a->tail_z_col >= 0 && a->ld_tail_in > a->tail_xy_col + 1 &&
                                          a->ld_tail_in > a->tail_z_col), TC_ERR_SHAPE, "tc_linear: bad tail columns");
    TC_REQUIRE(!a->tail_geom_out || aligned16(a->tail_geom_out), TC_ERR_ALIGN, "tc_linear: tail_geom_out must be 16-byte aligned");
    TC_REQUIRE(a->out_f32 != nullptr, TC_ERR_NULL, "tc_linear: a tail needs the fp32 output");
  }
  if (a->M == 0) return TC_OK;
  cudaStream_t s = as_stream(stream);
  if (linear_wide_supported(a)) return linear_wide_launch(a, s);
  if (linear_tc_supported(a)) return linear_tc_launch(a, s);
  return linear_simt_launch(a, s);
}

extern "C" int tc_attention_fwd(const tc_attention_args* a, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(a != nullptr, TC_ERR_NULL, "tc_attention_fwd: args is NULL");
  TC_REQUIRE(a->q && a->k && a->v && a->out, TC_ERR_NULL, "tc_attention_fwd: NULL tensor pointer");
  TC_REQUIRE(a->D == 32, TC_ERR_SHAPE, "tc_attention_fwd: head dim must be 32 (got %d)", a->D);
  TC_REQUIRE(a->B >= 0 && a->Lq >= 0 && a->Lk >= 0 && a->heads > 0 && a->B <= 65535 && a->heads <= 65535,
             TC_ERR_SHAPE, "tc_attention_fwd: bad shape");
  TC_REQUIRE(a->qkv_dtype == TC_F32 || a->qkv_dtype == TC_BF16 || a->qkv_dtype == TC_F16, TC_ERR_DTYPE,
             "tc_attention_fwd: bad qkv dtype");
  TC_REQUIRE(a->out_dtype == TC_F32 || a->out_dtype == TC_BF16 || a->out_dtype == TC_BF16X2, TC_ERR_DTYPE,
             "tc_attention_fwd: bad out dtype");
  TC_REQUIRE((a->geom == nullptr) == (a->key_xy == nullptr), TC_ERR_NULL,
             "tc_attention_fwd: geom and key_xy must be given together");
  const int64_t hd = (int64_t)a->heads * a->D;
  TC_REQUIRE(a->ldq >= hd && a->ldk >= hd && a->ldv >= hd && a->ldo >= hd, TC_ERR_SHAPE,
             "tc_attention_fwd: leading dimension smaller than heads*D");
  const int es = a->qkv_dtype == TC_F32 ? 4 : 2, eo = a->out_dtype == TC_F32 ? 4 : 2;
  TC_REQUIRE(aligned16(a->q) && aligned16(a->k) && aligned16(a->v) && aligned16(a->out) &&
                 (a->ldq * es) % 16 == 0 && (a->ldk * es) % 16 == 0 && (a->ldv * es) % 16 == 0 &&
                 (a->ldo * eo) % 16 == 0 && (a->q_batch_stride * es) % 16 == 0 && (a->k_batch_stride * es) % 16 == 0 &&
                 (a->v_batch_stride * es) % 16 == 0,
             TC_ERR_ALIGN, "tc_attention_fwd: q/k/v/out rows must be 16-byte aligned");
  TC_REQUIRE(a->algo >= TC_ATTN_AUTO && a->algo <= TC_ATTN_SPARSE, TC_ERR_SHAPE, "tc_attention_fwd: unknown algo %d", a->algo);
  TC_REQUIRE(!a->key_xy || (reinterpret_cast<uintptr_t>(a->key_xy) & 7u) == 0, TC_ERR_ALIGN,
             "tc_attention_fwd: key_xy must be 8-byte aligned");
  TC_REQUIRE(a->dropout_p >= 0.f && a->dropout_p < 1.f, TC_ERR_SHAPE, "tc_attention_fwd: dropout_p must be in [0, 1)");
  if (a->B == 0 || a->Lq == 0) return TC_OK;
  cudaStream_t s = as_stream(stream);
  if (a->attn_blocked || a->key_blocked) {      // generic dense masks: CUDA-core path only
    TC_REQUIRE(a->algo == TC_ATTN_AUTO || a->algo == TC_ATTN_SIMT, TC_ERR_SHAPE, "tc_attention_fwd: attn_blocked / key_blocked need the SIMT path");
    TC_REQUIRE(a->qkv_dtype != TC_F16 && a->out_dtype != TC_BF16X2, TC_ERR_DTYPE, "tc_attention_fwd: attn_blocked / key_blocked take fp32 / bf16 operands");
    return attention_simt_launch(a, s);
  }
  if (a->dropout_p > 0.f) {        // training variant: probability dropout lives in the fp32 kernels only
    TC_REQUIRE(a->algo != TC_ATTN_TENSOR && a->qkv_dtype != TC_F16 && a->out_dtype != TC_BF16X2, TC_ERR_DTYPE,
               "tc_attention_fwd: dropout is supported on the SIMT and sparse paths");
    if (a->algo != TC_ATTN_SIMT && attention_sparse_supported(a)) return attention_sparse_launch(a, s);
    return attention_simt_launch(a, s);
  }
  switch (a->algo) {
    case TC_ATTN_SPARSE:
      TC_REQUIRE(attention_sparse_supported(a), TC_ERR_SHAPE, "tc_attention_fwd: sparse path needs geom/key_xy and 8 heads x 32");
      return attention_sparse_launch(a, s);
    case TC_ATTN_TENSOR:
      TC_REQUIRE(attention_tc_supported(a), TC_ERR_DTYPE, "tc_attention_fwd: tensor-core path needs bf16 q/k/v/out");
      return attention_tc_launch(a, s);
    case TC_ATTN_SIMT:
      TC_REQUIRE(a->qkv_dtype != TC_F16 && a->out_dtype != TC_BF16X2, TC_ERR_DTYPE,
                 "tc_attention_fwd: the SIMT path takes fp32 / bf16 operands and outputs");
      return attention_simt_launch(a, s);
    default:
      break;
  }
  if (attention_sparse_supported(a)) return attention_sparse_launch(a, s);
  if (attention_tc_supported(a)) return attention_tc_launch(a, s);
  TC_REQUIRE(a->qkv_dtype != TC_F16 && a->out_dtype != TC_BF16X2, TC_ERR_DTYPE,
             "tc_attention_fwd: fp16 operands / split bf16 output need the tensor-core or sparse path");
  return attention_simt_launch(a, s);
}
