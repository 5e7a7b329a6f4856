// Box-regression losses on matched (prediction, target) pairs with their gradient w.r.t. the prediction:
//   mode 0  IOUloss "iou"   1 - iou^2                       yolov7/utils/boxes.py:125-151   (YOLOX head, yolox_head.py:134)
//   mode 1  IOUloss "giou"  1 - clamp(giou, -1, 1)          boxes.py:152-161
//   mode 2  IOUlossV6 giou  mode 3 diou  mode 4 ciou        boxes.py:666-752                 (YOLOv6 head; the north star's "CIoU")
// Boxes are (cx, cy, w, h).  The gradient is obtained by forward-mode differentiation with a 4-wide dual number, so the
// kernel text follows the reference formula line by line (including which epsilons are added where and the detached alpha
// of CIoU).  Sub-gradients of min / max / clamp match torch: ties between the two arguments split 0.5 / 0.5.
#include "host_common.cuh"
#include "sm100.cuh"

using namespace yb;

namespace {

struct D4 {  // value + partial derivatives w.r.t. (pred cx, cy, w, h)
  float v, g[4];
};
__device__ __forceinline__ D4 cst(float v) { return D4{v, {0.f, 0.f, 0.f, 0.f}}; }
__device__ __forceinline__ D4 var(float v, int i) {
  D4 r = cst(v);
  r.g[i] = 1.f;
  return r;
}
__device__ __forceinline__ D4 operator+(const D4& a, const D4& b) { return D4{a.v + b.v, {a.g[0] + b.g[0], a.g[1] + b.g[1], a.g[2] + b.g[2], a.g[3] + b.g[3]}}; }
__device__ __forceinline__ D4 operator-(const D4& a, const D4& b) { return D4{a.v - b.v, {a.g[0] - b.g[0], a.g[1] - b.g[1], a.g[2] - b.g[2], a.g[3] - b.g[3]}}; }
__device__ __forceinline__ D4 operator*(const D4& a, const D4& b) {
  return D4{a.v * b.v, {a.g[0] * b.v + a.v * b.g[0], a.g[1] * b.v + a.v * b.g[1], a.g[2] * b.v + a.v * b.g[2], a.g[3] * b.v + a.v * b.g[3]}};
}
__device__ __forceinline__ D4 operator/(const D4& a, const D4& b) {
  const float q = a.v / b.v, ib = 1.f / b.v;
  return D4{q, {(a.g[0] - q * b.g[0]) * ib, (a.g[1] - q * b.g[1]) * ib, (a.g[2] - q * b.g[2]) * ib, (a.g[3] - q * b.g[3]) * ib}};
}
__device__ __forceinline__ D4 operator+(const D4& a, float b) { D4 r = a; r.v += b; return r; }
__device__ __forceinline__ D4 operator-(const D4& a, float b) { D4 r = a; r.v -= b; return r; }
__device__ __forceinline__ D4 operator*(const D4& a, float b) { return D4{a.v * b, {a.g[0] * b, a.g[1] * b, a.g[2] * b, a.g[3] * b}}; }
__device__ __forceinline__ D4 operator-(float a, const D4& b) { return D4{a - b.v, {-b.g[0], -b.g[1], -b.g[2], -b.g[3]}}; }
__device__ __forceinline__ D4 mix(const D4& a, const D4& b, float wa) {  // wa*a + (1-wa)*b on the derivatives
  const float wb = 1.f - wa;
  return D4{wa >= 0.5f ? a.v : b.v, {wa * a.g[0] + wb * b.g[0], wa * a.g[1] + wb * b.g[1], wa * a.g[2] + wb * b.g[2], wa * a.g[3] + wb * b.g[3]}};
}
__device__ __forceinline__ D4 dmax(const D4& a, const D4& b) { return a.v > b.v ? a : (a.v < b.v ? b : mix(a, b, 0.5f)); }
__device__ __forceinline__ D4 dmin(const D4& a, const D4& b) { return a.v < b.v ? a : (a.v > b.v ? b : mix(a, b, 0.5f)); }
__device__ __forceinline__ D4 clamp_min(const D4& a, float lo) { return a.v >= lo ? a : cst(lo); }  // torch.clamp: gradient 1 at the bound
__device__ __forceinline__ D4 clamp(const D4& a, float lo, float hi) { return (a.v >= lo && a.v <= hi) ? a : cst(a.v < lo ? lo : hi); }
__device__ __forceinline__ D4 datan(const D4& a) {
  const float d = 1.f / (1.f + a.v * a.v);
  return D4{atanf(a.v), {a.g[0] * d, a.g[1] * d, a.g[2] * d, a.g[3] * d}};
}

__global__ void iou_loss_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, int n, int mode, float* __restrict__ loss,
                                float* __restrict__ dpred) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const D4 px = var(pred[4 * i], 0), py = var(pred[4 * i + 1], 1), pw = var(pred[4 * i + 2], 2), ph = var(pred[4 * i + 3], 3);
  const D4 tx = cst(tgt[4 * i]), ty = cst(tgt[4 * i + 1]), tw = cst(tgt[4 * i + 2]), th = cst(tgt[4 * i + 3]);
  const D4 p_l = px - pw * 0.5f, p_r = px + pw * 0.5f, p_t = py - ph * 0.5f, p_b = py + ph * 0.5f;
  const D4 t_l = tx - tw * 0.5f, t_r = tx + tw * 0.5f, t_t = ty - th * 0.5f, t_b = ty + th * 0.5f;
  D4 l;
  if (mode <= 1) {
    // IOUloss (boxes.py:131-161)
    const D4 tlx = dmax(p_l, t_l), tly = dmax(p_t, t_t), brx = dmin(p_r, t_r), bry = dmin(p_b, t_b);
    const float en = (tlx.v < brx.v && tly.v < bry.v) ? 1.f : 0.f;
    const D4 area_i = (brx - tlx) * (bry - tly) * en;
    const D4 area_p = pw * ph, area_g = tw * th;
    const D4 iou = area_i / (area_p + area_g - area_i + 1e-16f);
    if (mode == 0) {
      l = 1.f - iou * iou;
    } else {
      const D4 c_tlx = dmin(p_l, t_l), c_tly = dmin(p_t, t_t), c_brx = dmax(p_r, t_r), c_bry = dmax(p_b, t_b);
      const D4 area_c = (c_brx - c_tlx) * (c_bry - c_tly);
      const D4 giou = iou - (area_c - area_i) / clamp_min(area_c, 1e-16f);
      l = 1.f - clamp(giou, -1.f, 1.f);
    }
  } else {
    // IOUlossV6, box_format "xywh" (boxes.py:696-733), eps = 1e-7
    const float eps = 1e-7f;
    const D4 inter = clamp_min(dmin(p_r, t_r) - dmax(p_l, t_l), 0.f) * clamp_min(dmin(p_b, t_b) - dmax(p_t, t_t), 0.f);
    const D4 w1 = p_r - p_l, h1 = p_b - p_t + eps, w2 = t_r - t_l, h2 = t_b - t_t + eps;
    const D4 uni = w1 * h1 + w2 * h2 - inter + eps;
    D4 iou = inter / uni;
    const D4 cw = dmax(p_r, t_r) - dmin(p_l, t_l), ch = dmax(p_b, t_b) - dmin(p_t, t_t);
    if (mode == 2) {
      const D4 c_area = cw * ch + eps;
      iou = iou - (c_area - uni) / c_area;
    } else {
      const D4 c2 = cw * cw + ch * ch + eps;
      const D4 dx = t_l + t_r - p_l - p_r, dy = t_t + t_b - p_t - p_b;
      const D4 rho2 = (dx * dx + dy * dy) * 0.25f;
      if (mode == 3) {
        iou = iou - rho2 / c2;
      } else {
        const D4 da = datan(w2 / h2) - datan(w1 / h1);
        const D4 v = da * da * (4.f / (3.14159265358979323846f * 3.14159265358979323846f));
        const float alpha = v.v / (v.v - iou.v + (1.f + eps));  // computed under torch.no_grad(): a constant
        iou = iou - (rho2 / c2 + v * alpha);
      }
    }
    l = 1.f - iou;
  }
  loss[i] = l.v;
  if (dpred) {
    dpred[4 * i] = l.g[0];
    dpred[4 * i + 1] = l.g[1];
    dpred[4 * i + 2] = l.g[2];
    dpred[4 * i + 3] = l.g[3];
  }
}

}  // namespace

extern "C" int yb200_iou_loss(const float* pred_cxcywh, const float* target_cxcywh, int n, int mode, float* loss, float* dloss_dpred,
                              void* stream) {
  YB_REQUIRE(pred_cxcywh && target_cxcywh && loss, YB200_ERR_INVALID, "iou_loss: null pointer");
  YB_REQUIRE(n >= 0 && mode >= 0 && mode <= 4, YB200_ERR_INVALID, "iou_loss: n=%d mode=%d", n, mode);
  if (n == 0) return 0;
  launch_k(iou_loss_kernel, ceil_div(n, 128), 128, 0, as_stream(stream), pred_cxcywh, target_cxcywh, n, mode, loss, dloss_dpred);
  YB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
