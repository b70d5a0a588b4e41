// N4: the loss side of the TransCAR head on the device.
//   tc_match_cost   Hungarian cost matrix of every (output layer, sample) problem in one launch:
//                   FocalLossCost (mmdet) + BBox3DL1Cost (match_cost.py:15-26) against normalize_bbox(gt) (util.py:4-24).
//                   Replaces hungarian_assigner_3d.py:106-115 (two dozen ATen launches per problem).
//   tc_detr_loss    sigmoid focal loss + code-weighted L1 loss of detr3d_head.py:849-917 for all layers at once, fused with
//                   their analytic gradients (d loss / d logits, d loss / d box codes): the backward pass of the head starts
//                   from this kernel's outputs, no autograd graph over the loss.
// The assignment itself (scipy.optimize.linear_sum_assignment, hungarian_assigner_3d.py:117-124) stays on the host like in the
// reference; transcar_b200/loss.py overlaps it with the device.
#include "tc_common.cuh"

namespace tc {
namespace {

__device__ __forceinline__ float softplus_f32(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

// util.py:4-24  (cx, cy, cz, w, l, h, rot, vx, vy) -> (cx, cy, log w, log l, cz, log h, sin, cos, vx, vy)
__device__ __forceinline__ void normalize_box(const float* g, float (&t)[10]) {
  t[0] = g[0]; t[1] = g[1]; t[2] = logf(g[3]); t[3] = logf(g[4]); t[4] = g[2]; t[5] = logf(g[5]);
  t[6] = sinf(g[6]); t[7] = cosf(g[6]); t[8] = g[7]; t[9] = g[8];
}

__global__ void __launch_bounds__(256) match_cost_kernel(const tc_match_cost_args a) {
  const int p = blockIdx.y;                          // problem = layer * B + sample
  const int b = p % a.B;
  const int g0 = a.gt_offsets[b], G = a.gt_offsets[b + 1] - g0;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.Q * a.Gmax) return;
  const int q = (int)(idx / a.Gmax), g = (int)(idx % a.Gmax);
  float* out = a.cost + ((long long)p * a.Q + q) * a.Gmax + g;
  if (g >= G) { *out = INFINITY; return; }
  const int label = a.gt_labels[g0 + g];
  const float x = a.cls[((long long)p * a.Q + q) * a.classes + label];
  // mmdet FocalLossCost: sigmoid, eps inside the logs
  const float s = sigmoid_f32(x);
  const float neg = -logf(1.0f - s + a.eps) * (1.0f - a.alpha) * powf(s, a.gamma);
  const float pos = -logf(s + a.eps) * a.alpha * powf(1.0f - s, a.gamma);
  float t[10];
  normalize_box(a.gt_boxes + (long long)(g0 + g) * 9, t);
  const float* bp = a.bbox + ((long long)p * a.Q + q) * 10;
  float l1 = 0.f;
#pragma unroll
  for (int j = 0; j < 10; ++j) l1 += fabsf(bp[j] - t[j]);
  *out = (pos - neg) * a.cls_weight + l1 * a.reg_weight;
}

__global__ void __launch_bounds__(256) detr_loss_kernel(const tc_detr_loss_args a) {
  __shared__ float s_cls[8], s_box[8];
  const int p = blockIdx.y, layer = p / a.B;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  float lc = 0.f, lb = 0.f;
  if (q < a.Q) {
    const long long row = (long long)p * a.Q + q;
    const int gi = a.assigned[row];                  // -1 = background, else global ground-truth index
    const int label = gi >= 0 ? a.gt_labels[gi] : -1;
    const float cls_scale = a.loss_cls_weight / a.cls_avg[layer];
    for (int c = 0; c < a.classes; ++c) {
      const float x = a.cls[row * a.classes + c];
      const float pr = sigmoid_f32(x);
      float loss, grad;
      if (c == label) {                              // target 1: alpha (1-p)^gamma * -log p
        const float w = a.alpha * powf(1.0f - pr, a.gamma);
        loss = w * softplus_f32(-x);
        grad = w * (a.gamma * pr * (-softplus_f32(-x)) - (1.0f - pr));
      } else {                                       // target 0: (1-alpha) p^gamma * -log(1-p)
        const float w = (1.0f - a.alpha) * powf(pr, a.gamma);
        loss = w * softplus_f32(x);
        grad = w * (pr + a.gamma * (1.0f - pr) * softplus_f32(x));
      }
      if (isnan(loss)) { loss = 0.f; grad = 0.f; }
      lc += loss * cls_scale;
      if (a.d_cls) a.d_cls[row * a.classes + c] = grad * cls_scale;
    }
    float dbox[10] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (gi >= 0) {
      float t[10];
      normalize_box(a.gt_boxes + (long long)gi * 9, t);
      bool finite = true;
#pragma unroll
      for (int j = 0; j < 10; ++j) finite = finite && isfinite(t[j]);
      if (finite) {                                  // H:899 isnotnan filter
        const float box_scale = a.loss_bbox_weight / a.pos_avg[layer];
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          const float d = a.bbox[row * 10 + j] - t[j];
          const float w = a.code_weights[j] * box_scale;
          lb += fabsf(d) * w;
          dbox[j] = d > 0.f ? w : (d < 0.f ? -w : 0.f);
        }
      }
    }
    if (a.d_bbox) {
#pragma unroll
      for (int j = 0; j < 10; ++j) a.d_bbox[row * 10 + j] = dbox[j];
    }
  }
  lc = warp_sum(lc);
  lb = warp_sum(lb);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_cls[warp] = lc; s_box[warp] = lb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float c = 0.f, bx = 0.f;
    for (int w = 0; w < 8; ++w) { c += s_cls[w]; bx += s_box[w]; }
    atomicAdd(a.loss_cls + layer, c);
    atomicAdd(a.loss_bbox + layer, bx);
  }
}

}  // namespace
}  // namespace tc

extern "C" int tc_match_cost(const tc_match_cost_args* a, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(a != nullptr, TC_ERR_NULL, "tc_match_cost: args is NULL");
  TC_REQUIRE(a->layers >= 0 && a->B >= 0 && a->Q >= 0 && a->classes > 0 && a->Gmax >= 0, TC_ERR_SHAPE, "tc_match_cost: bad shape");
  if (a->layers == 0 || a->B == 0 || a->Q == 0 || a->Gmax == 0) return TC_OK;
  TC_REQUIRE(a->cls && a->bbox && a->gt_boxes && a->gt_labels && a->gt_offsets && a->cost, TC_ERR_NULL, "tc_match_cost: NULL pointer");
  TC_REQUIRE(a->layers * a->B <= 65535, TC_ERR_SHAPE, "tc_match_cost: too many problems");
  const long long per = (long long)a->Q * a->Gmax;
  match_cost_kernel<<<dim3((unsigned)((per + 255) / 256), a->layers * a->B), 256, 0, as_stream(stream)>>>(*a);
  count_launch();
  return check_launch("tc_match_cost");
}

extern "C" int tc_detr_loss(const tc_detr_loss_args* a, tc_stream_t stream) {
  using namespace tc;
  TC_REQUIRE(a != nullptr, TC_ERR_NULL, "tc_detr_loss: args is NULL");
  TC_REQUIRE(a->layers >= 0 && a->B >= 0 && a->Q >= 0 && a->classes > 0, TC_ERR_SHAPE, "tc_detr_loss: bad shape");
  if (a->layers == 0 || a->B == 0 || a->Q == 0) return TC_OK;
  TC_REQUIRE(a->cls && a->bbox && a->assigned && a->code_weights && a->cls_avg && a->pos_avg && a->loss_cls && a->loss_bbox,
             TC_ERR_NULL, "tc_detr_loss: NULL pointer");
  TC_REQUIRE(a->layers * a->B <= 65535, TC_ERR_SHAPE, "tc_detr_loss: too many problems");
  detr_loss_kernel<<<dim3((a->Q + 255) / 256, a->layers * a->B), 256, 0, as_stream(stream)>>>(*a);
  count_launch();
  return check_launch("tc_detr_loss");
}
