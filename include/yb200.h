/* yb200 -- C ABI of the B200 (sm_100a) YOLOX hot path.
 *
 * The reference (lucasjinreal/yolov7_d2) has no FFI of its own: every entry point below replaces a
 * PyTorch / torchvision call the reference makes on its hot path (file:line cited per function, paths
 * relative to the reference tree).  All pointers are DEVICE pointers unless stated otherwise, all tensors
 * are caller-allocated, every call is asynchronous on `stream` (a cudaStream_t passed as void*), performs no
 * host synchronisation and returns 0 on success or a negative yb200_status.  No global state except a
 * lazily resolved driver entry point (cuTensorMapEncodeTiled) and the per-device SM count.
 *
 * Activation layout: NHWC bf16 with a channel pitch, so that channel slices of a concat buffer are views:
 *   element (n,y,x,c) of a view = ptr[((n*h + y)*w + x)*c_pitch + c_off + c]
 */
#ifndef YB200_H_
#define YB200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YB200_VERSION 100

typedef enum {
  YB200_OK = 0,
  YB200_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, misaligned pitch) */
  YB200_ERR_UNSUPPORTED = -2, /* shape outside what the kernels implement */
  YB200_ERR_CUDA = -3,        /* a CUDA runtime / driver call failed; see yb200_last_error() */
} yb200_status;

typedef struct {
  void* ptr;        /* base of the underlying NHWC bf16 buffer (16-byte aligned) */
  int32_t n, h, w;  /* logical extents */
  int32_t c;        /* channels of this view (multiple of 8) */
  int32_t c_pitch;  /* channels of the underlying buffer (multiple of 8) */
  int32_t c_off;    /* first channel of the view inside the buffer (multiple of 8) */
} yb200_act;

int yb200_version(void);
const char* yb200_last_error(void);

/* ---- weights ------------------------------------------------------------------------------------- */
/* nn.Conv2d.weight (fp32 OIHW, wrappers.py:67-75) -> the two bf16 GEMM operands used by the kernels:
 *   w_fwd  [cout_pad][k*k][cin_pad]  (forward / weight-gradient order),  zero padded
 *   w_dgrad[cin_pad][k*k][cout_pad]  (data-gradient order); may be NULL.                                 */
int yb200_pack_conv_weight(const float* w_oihw, int cout, int cin, int ksize, int cout_pad, int cin_pad,
                           void* w_fwd, void* w_dgrad, void* stream);

/* ---- convolution (implicit GEMM on tcgen05) ------------------------------------------------------ */
/* z = conv2d(x, w) without bias, padding (k-1)/2 -- BaseConv.conv, wrappers.py:67-80.
 * ksize in {1,3}, stride in {1,2} (stride 2 only with ksize 3).  z is the pre-BatchNorm output, stored as **fp16**
 * (same 2-byte NHWC view; BatchNorm's mean subtraction makes this tensor the precision-critical one).
 * If stat_sum/stat_sqsum are non-NULL the per-channel sum and sum of squares of the *stored* z are
 * accumulated into them (fp64, must be zeroed by the caller) -- the batch statistics nn.BatchNorm2d
 * (wrappers.py:76) computes in training mode.                                                          */
int yb200_conv2d_fwd(const yb200_act* x, const void* w_fwd, const yb200_act* z, int ksize, int stride,
                     double* stat_sum, double* stat_sqsum, void* stream);

/* Eval-mode BaseConv in one kernel: out = SiLU(conv2d(x, w)*scale + shift) [+ residual], bf16 -- nn.Conv2d + nn.BatchNorm2d
 * (running statistics) + nn.SiLU of wrappers.py:60-83 with the BatchNorm folded as in utils/checkpoint.py:11-43
 * (scale / shift from yb200_bn_eval_affine); the Bottleneck shortcut (wrappers.py:119-123) is added after the activation. */
int yb200_conv2d_bn_silu_fwd(const yb200_act* x, const void* w_fwd, const float* scale, const float* shift,
                             const yb200_act* residual, const yb200_act* out, int ksize, int stride, void* stream);

/* yb200_conv2d_fwd on a PIXEL-GROUPED view: when g horizontally adjacent pixels are viewed as one pixel of g*C channels (same memory), a
 * convolution becomes a convolution of the grouped tensors with an expanded weight matrix; output column k is channel k % stat_fold of one of
 * the g pixels, and the BatchNorm sums accumulate per channel.  For ksize 3 the weights MUST be such an expansion: the kernel skips the products of
 * the left / right neighbour group that the expansion makes zero (all but its last / first pixel).  Used for the stem (12 -> 32 channels at 320x320: 32-byte pixel rows are bound by
 * the TMA row rate; grouped by 4 they are 128-byte rows), `BaseConv` of `backbone.stem.conv` (darknetx.py:117, wrappers.py:60-80).            */
int yb200_conv2d_fwd_fold(const yb200_act* x, const void* w_fwd, const yb200_act* z, int ksize, int stride,
                          double* stat_sum, double* stat_sqsum, int stat_fold, void* stream);

/* out[n, a_off + y*w + x, c_off + c] = conv1x1(x, w)[n,y,x,c] + bias[c] in fp32 -- the prediction convs
 * yolox_head.py:103-129 fused with the cat/flatten/permute of yolox_head.py:175,238-244.
 * out is [n][a_total][c_total] fp32.                                                                    */
int yb200_conv1x1_bias_f32(const yb200_act* x, const void* w_fwd, const float* bias, int cout, float* out,
                           int a_total, int a_off, int c_total, int c_off, void* stream);

/* dx = conv_transpose(dz, w) [+ addend] -- autograd of the convolution w.r.t. its input.  dx/addend have the
 * input's shape, dz the output's.  addend may be NULL.                                                   */
int yb200_conv2d_dgrad(const yb200_act* dz, const void* w_dgrad, const yb200_act* dx, const yb200_act* addend,
                       int ksize, int stride, void* stream);
/* The same data gradient, whose epilogue ALSO performs the reduction pass of BatchNorm+SiLU backward for the layer(s) that
 * produced the activation whose gradient dx is (autograd of wrappers.py:76-80): for every segment, channels
 * [dx_c_begin, dx_c_begin + z.c) of dx are the final gradient of a = SiLU(z*scale + shift); the kernel accumulates
 * sum_du[c] += sum da*SiLU'(u), sum_duz[c] += sum da*SiLU'(u)*z  (fp64, c relative to the segment) while da is still in
 * registers, reading z (fp16, dx's pixel grid) once.  Feed the sums to yb200_bn_silu_bwd_apply.  dx must be the LAST
 * contribution to that gradient (addend = the earlier ones); gradient tensors of >= 256 channels are not supported.      */
typedef struct yb200_bnbwd_seg {
  yb200_act z;          /* fp16 pre-BatchNorm output of the producer (n, h, w as dx) */
  int32_t dx_c_begin;   /* first channel of dx covered by this segment (multiple of 32) */
  const float* scale;   /* per channel of the segment */
  const float* shift;
  double* sum_du;
  double* sum_duz;
} yb200_bnbwd_seg;
int yb200_conv2d_dgrad_bnbwd(const yb200_act* dz, const void* w_dgrad, const yb200_act* dx, const yb200_act* addend,
                             int ksize, int stride, int num_segments, const yb200_bnbwd_seg* segments, void* stream);

/* grad_oihw (+)= d loss / d weight, fp32 [cout][cin_real][k][k] -- autograd of the convolution w.r.t. its weight.
 * workspace: at least yb200_conv2d_wgrad_workspace() bytes.                                              */
int64_t yb200_conv2d_wgrad_workspace(const yb200_act* x, const yb200_act* dz, int ksize, int stride);
int yb200_conv2d_wgrad(const yb200_act* x, const yb200_act* dz, int ksize, int stride, int cin_real,
                       float* grad_oihw, int accumulate, void* workspace, int64_t workspace_bytes, void* stream);
/* The same for the PIXEL-GROUPED stem (x, dz = the grouped views of yb200_conv2d_fwd_fold, `group` pixels per group): grad_oihw is the gradient of
 * the EXPANDED weight matrix, valid at the positions the expansion fills (the caller folds them onto the [cout, cin, 3, 3] parameter); positions
 * that the expansion leaves zero are not computed (side taps multiply one neighbour pixel instead of the whole group).                      */
int yb200_conv2d_wgrad_grouped(const yb200_act* x, const yb200_act* dz, int ksize, int stride, int cin_real, int group,
                               float* grad_oihw, int accumulate, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- preprocessing ------------------------------------------------------------------------------- */
/* uint8 NCHW images -> Focus (space-to-depth) NHWC bf16 [n, h/2, w/2, 16]: channel = patch*3 + rgb with patch order
 * (top-left, bottom-left, top-right, bottom-right) as wrappers.py:210-220, channels 12..15 zero.  No mean/std
 * normalisation (yolox.py:98).  hw_valid (device int32 [n][2], may be NULL) gives each image's true (h, w); pixels
 * beyond it read as pad_value = MODEL.PADDED_VALUE 114 (yolox.py:100-101, detectron2 ImageList.from_tensors).   */
int yb200_preprocess_focus(const uint8_t* images_nchw, int n, int h, int w, const int32_t* hw_valid, float pad_value,
                           const yb200_act* out, void* stream);

/* ---- BatchNorm + SiLU (nn.BatchNorm2d + nn.SiLU of BaseConv, wrappers.py:76-80; eps/momentum yolox.py:85-90) ---- */
/* Training statistics -> per-channel affine: scale = gamma*invstd, shift = beta - mean*scale (biased variance);
 * updates running_mean / running_var (unbiased variance, momentum) and num_batches_tracked in place when given;
 * zeroes stat_sum / stat_sqsum for the next step.                                                          */
int yb200_bn_finalize(double* stat_sum, double* stat_sqsum, int c, int64_t count, const float* gamma, const float* beta,
                      float eps, float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                      float* scale, float* shift, float* save_mean, float* save_invstd, void* stream);
/* Eval mode: scale/shift from running statistics (the folding of utils/checkpoint.py:11-43 as an epilogue).   */
int yb200_bn_eval_affine(int c, const float* gamma, const float* beta, const float* running_mean,
                         const float* running_var, float eps, float* scale, float* shift, void* stream);
/* z is the fp16 tensor written by yb200_conv2d_fwd; every other activation / gradient view is bf16.
 * out = SiLU(z*scale + shift) [+ residual]  (Bottleneck shortcut, wrappers.py:119-123); when out_up2x is given the
 * result is also written nearest-upsampled x2 into that view (nn.Upsample + torch.cat of yolo_pafpn.py:96-102). */
int yb200_bn_apply_silu(const yb200_act* z, const float* scale, const float* shift, const yb200_act* residual,
                        const yb200_act* out, const yb200_act* out_up2x, void* stream);
/* yb200_bn_finalize + yb200_bn_apply_silu in ONE launch (training): every block derives scale / shift of its channels from the batch
 * sums, block 0 publishes scale / shift / save_mean / save_invstd (read by the backward pass) and updates the running statistics.
 * All pointers are per-channel arrays of this view's channels.  stat_sum / stat_sqsum are NOT cleared (clear them once per step).      */
int yb200_bn_train_apply_silu(const yb200_act* z, const double* stat_sum, const double* stat_sqsum, int64_t count,
                              const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                              float* running_var, float* scale, float* shift, float* save_mean, float* save_invstd,
                              const yb200_act* residual, const yb200_act* out, const yb200_act* out_up2x, void* stream);
/* Backward of SiLU(BN(z)) in training mode.  The incoming gradient is da [+ da2] [+ 2x2 sum-pool of da_up2x]
 * (fan-out of the activation / backward of the upsample).  Writes dz (bf16) and dgamma / dbeta (fp32, optionally
 * accumulated).  acc_dgamma / acc_dbeta: fp64 scratch [c], must be zero on entry, are zero on exit.          */
int yb200_bn_silu_bwd(const yb200_act* z, const yb200_act* da, const yb200_act* da2, const yb200_act* da_up2x,
                      const float* scale, const float* shift, const float* save_mean, const float* save_invstd,
                      double* acc_dgamma, double* acc_dbeta, const yb200_act* dz, float* dgamma, float* dbeta,
                      int accumulate, void* stream);
/* Second half of the same backward when the reduction already happened in the epilogue of yb200_conv2d_dgrad_bnbwd:
 * sum_duz / sum_du hold S2 = sum du*z and S1 = sum du (fp64 [c]; zeroed on exit).  dgamma = invstd*(S2 - mean*S1),
 * dbeta = S1, dz = scale*du - (scale*invstd*dgamma/M)*(z - mean) - scale*dbeta/M.                                      */
int yb200_bn_silu_bwd_apply(const yb200_act* z, const yb200_act* da, const float* scale, const float* shift,
                            const float* save_mean, const float* save_invstd, double* sum_duz, double* sum_du,
                            const yb200_act* dz, float* dgamma, float* dbeta, int accumulate, void* stream);

/* yb200_bn_silu_bwd / _apply with dgamma == dbeta == NULL leave their sums in the fp64 accumulators; this call converts the accumulators of a
 * run of c channels (several layers) in one launch: grad_base[gamma_off[i]] (+)= dgamma_i, grad_base[beta_off[i]] (+)= dbeta_i, accumulators
 * zeroed.  raw_sums (may be NULL): per channel, 1 where the accumulators hold S2 / S1 of yb200_conv2d_dgrad_bnbwd (then save_mean / save_invstd
 * are read).  Replaces the tail of torch's batch_norm backward for every BaseConv of a plan (wrappers.py:60-80).                        */
int yb200_bn_param_grads(double* acc_dgamma, double* acc_dbeta, int c, const int32_t* gamma_off, const int32_t* beta_off,
                         float* grad_base, const float* save_mean, const float* save_invstd, const uint8_t* raw_sums,
                         int accumulate, void* stream);

/* ---- SPP / concat helpers ------------------------------------------------------------------------ */
/* nn.MaxPool2d(k, 1, k//2) for k = 5, 9, 13 written into three channel slices (SPPBottleneck, wrappers.py:150-160).
 * argmax (may be NULL): uint8 [3][n][h][w][c] window offsets for the backward.                                */
int yb200_spp_pool(const yb200_act* x, const yb200_act* o5, const yb200_act* o9, const yb200_act* o13, uint8_t* argmax,
                   void* stream);
/* dx = d0 + maxpool-backward(d5, d9, d13);  scratch: fp32 [n*h*w*c].                                           */
int yb200_spp_pool_bwd(const yb200_act* d0, const yb200_act* d5, const yb200_act* d9, const yb200_act* d13,
                       const uint8_t* argmax, float* scratch, const yb200_act* dx, void* stream);
int yb200_copy_view(const yb200_act* src, const yb200_act* dst, void* stream);

/* ---- YOLOX head tail: decode, SimOTA, losses ----------------------------------------------------- */
/* level_hw_stride: HOST int32 [num_levels][3] = (h, w, stride) of each FPN level, anchors ordered level by level,
 * row-major (yolox_head.py:226-245).  outputs: [batch][num_anchors][5+C] fp32 = (x, y, w, h, obj, cls...).     */
/* In place: xy = (xy + grid)*stride, wh = exp(wh)*stride (yolox_head.py:238-244); eval_mode also applies sigmoid to
 * obj / cls (yolox_head.py:209-211, 247-272).                                                                  */
int yb200_yolox_decode(float* outputs, int batch, int num_anchors, int channels, const int32_t* level_hw_stride,
                       int num_levels, int eval_mode, void* stream);
/* Training decode that also keeps the RAW regression outputs: raw_reg [batch][num_anchors][4] fp32 (16-byte aligned) = `origin_preds` of the
 * L1 branch (`use_l1`, yolox_head.py:186-195).                                                                                           */
int yb200_yolox_decode_keep_raw(float* outputs, int batch, int num_anchors, int channels, const int32_t* level_hw_stride,
                                int num_levels, float* raw_reg, void* stream);
/* SimOTA dynamic-k assignment for the whole batch (get_assignments + get_in_boxes_info + dynamic_k_matching,
 * yolox_head.py:450-669) on DECODED outputs and labels [batch][max_gt][5] = (cls, cx, cy, w, h), zero padded.
 * Per anchor: fg_mask (u8), matched_gt (-1 when background), matched_iou, matched_cls; per image num_gt / num_fg;
 * totals[0] = number of foreground anchors in the batch, totals[1] = number of gts.                            */
int64_t yb200_simota_workspace(int batch, int num_anchors);
int yb200_simota_assign(const float* outputs, const float* labels, int batch, int num_anchors, int channels, int max_gt,
                        const int32_t* level_hw_stride, int num_levels, void* workspace, int32_t* num_gt,
                        uint8_t* fg_mask, int32_t* matched_gt, float* matched_iou, int32_t* matched_cls,
                        int32_t* num_fg_img, int32_t* totals, void* stream);
/* Losses of get_losses (yolox_head.py:412-441) and/or their gradient w.r.t. the RAW (pre-decode) head outputs.
 *   losses6 (device float[6], may be NULL): total, 5*iou, obj, cls, l1 (= 0), num_fg/num_gt.
 *   weights3 (device float[3], NULL = no gradient): d objective / d (loss_iou, loss_obj, loss_cls)  (5, 1, 1 for `total`).
 *   d_cls[l] / d_regobj[l] (HOST arrays of device pointers, one per level): bf16 NHWC [batch][h][w][C] / [..][16]
 *   (reg 0-3, obj 4, zero padding) -- the dz tensors of the prediction convs.  d_dense: optional fp32 [batch][A][5+C].
 *   bias_acc: optional fp64 [num_levels][5+C] accumulators (zeroed by the caller) of the summed gradients.
 *   loss_acc3: fp64 scratch [3], zero on entry and exit.                                                        */
int yb200_yolox_loss(const float* outputs, const float* labels, int batch, int num_anchors, int channels, int max_gt,
                     const int32_t* level_hw_stride, int num_levels, const uint8_t* fg_mask, const int32_t* matched_gt,
                     const float* matched_iou, const int32_t* matched_cls, const int32_t* totals, const float* weights3,
                     double* loss_acc3, float* losses6, void* const* d_cls, void* const* d_regobj, float* d_dense,
                     double* bias_acc, void* stream);
/* The same with the L1 branch (`use_l1`, yolox_head.py:389-429, 443-448): loss_l1 = sum |raw_reg[fg] - l1_target| / num_fg with
 * l1_target = (gt_xy / stride - grid, log(gt_wh / stride + 1e-8)); total = 5*iou + obj + cls + l1, losses6[4] = l1; weights4[3] = d objective /
 * d loss_l1; the gradient sign(raw - target) * weight / num_fg is added to the regression gradients.  loss_acc4: fp64 scratch [4].          */
int yb200_yolox_loss_l1(const float* outputs, const float* raw_reg, const float* labels, int batch, int num_anchors, int channels,
                        int max_gt, const int32_t* level_hw_stride, int num_levels, const uint8_t* fg_mask,
                        const int32_t* matched_gt, const float* matched_iou, const int32_t* matched_cls, const int32_t* totals,
                        const float* weights4, double* loss_acc4, float* losses6, void* const* d_cls, void* const* d_regobj,
                        float* d_dense, double* bias_acc, void* stream);
/* bias gradients of reg_preds / obj_preds / cls_preds of one level out of bias_acc (which is re-zeroed).          */
int yb200_head_bias_grad(double* bias_acc, int num_levels, int channels, int level, float* grad_reg_bias4,
                         float* grad_obj_bias1, float* grad_cls_bias, int accumulate, void* stream);

/* All convolution weights of a plan repacked in ONE launch (same result as n calls of yb200_pack_conv_weight).  table / prefix live in
 * DEVICE memory and are built once per plan: prefix[i] = sum of cout_pad*k*k*cin_pad over layers < i, prefix[n] = total.               */
typedef struct yb200_pack_desc {
  const float* w_oihw;
  void* w_fwd;      /* [cout_pad][k*k][cin_pad] bf16, may be NULL */
  void* w_dgrad;    /* [cin_pad][k*k][cout_pad] bf16, may be NULL */
  int32_t cout, cin, ksize, cout_pad, cin_pad, reserved;
} yb200_pack_desc;
int yb200_pack_conv_weights_batched(const yb200_pack_desc* table_dev, const int64_t* prefix_dev, int n, int64_t total, void* stream);

/* ---- strict mode: fp32-grade forward on the same tensor-core kernels -------------------------------- */
/* SPLIT storage: an activation a is the sum of `planes` bf16 values a0 = bf16(a), a1 = bf16(a - a0) [, a2 = bf16(a - a0 - a1)]:
 * 16 significant bits with planes = 2, the full 24 bits of fp32 with planes = 3.  All planes live in one NHWC bf16 buffer,
 * lo_delta channels apart, so a yb200_act describes plane 0 and (view, lo_delta, planes) the value.  The convolution keeps
 * every partial product a_i * w_j with i + j < planes (3 / 6 taps per spatial tap of the SAME tcgen05 implicit GEMM, one
 * fp32 accumulator); the pre-BatchNorm output z is fp32 NHWC [n][h/stride][w/stride][z_pitch], channels [z_off, z_off+cout).
 * Used to check the reference's fp32 logits / losses to 1e-3 (tests/test_strict_gpu.py); forward only.
 * yb200_pack_conv_weight_split: fp32 OIHW -> [cout_pad][planes][k*k][cin_pad], `nn.Conv2d.weight` of wrappers.py:67-75.
 * yb200_conv2d_fwd_split: `BaseConv.conv` (wrappers.py:79).  yb200_conv1x1_bias_f32_split: the prediction convolutions
 * (yolox_head.py:103-129,175), same output geometry as yb200_conv1x1_bias_f32.                                           */
int yb200_pack_conv_weight_split(const float* w_oihw, int cout, int cin, int ksize, int cout_pad, int cin_pad, int planes,
                                 void* w_split, void* stream);
int yb200_conv2d_fwd_split(const yb200_act* x, int lo_delta, int planes, const void* w_split, int cout, int ksize,
                           int stride, float* z, int z_pitch, int z_off, void* stream);
int yb200_conv1x1_bias_f32_split(const yb200_act* x, int lo_delta, int planes, const void* w_split, const float* bias,
                                 int cout, float* out, int a_total, int a_off, int c_total, int c_off, void* stream);
/* BatchNorm batch statistics of an fp32 slice (sum, sum of squares in fp64; feed yb200_bn_finalize), then
 * out = SiLU(z*scale + shift) [+ residual] as split planes, optionally also 2x nearest-upsampled (wrappers.py:76-80,
 * 119-123; yolo_pafpn.py:96-102); SPP max-pools 5/9/13 on split values (wrappers.py:150-160).                        */
int yb200_strict_bn_stats(const float* z, int64_t npix, int z_pitch, int z_off, int c, double* stat_sum,
                          double* stat_sqsum, void* stream);
int yb200_strict_bn_apply_silu(const float* z, int z_pitch, int z_off, const float* scale, const float* shift,
                               const yb200_act* residual, int residual_lo, const yb200_act* out, int out_lo,
                               const yb200_act* out_up2x, int up_lo, int planes, void* stream);
int yb200_strict_spp_pool(const yb200_act* x, const yb200_act* o5, const yb200_act* o9, const yb200_act* o13,
                          int lo_delta, int planes, void* stream);

/* ---- post-processing ----------------------------------------------------------------------------- */
/* `postprocess` (boxes.py:171-210) for the whole batch: prediction [batch][A][5+C] = (cx, cy, w, h, obj, cls...) with
 * probabilities already applied.  Per image: class_conf / class_pred = max / first argmax over classes; candidates have
 * obj*class_conf >= conf_thre; per-class NMS (torchvision batched_nms, "vanilla" semantics: stable descending score
 * order, suppress when IoU > nms_thre on un-offset fp32 xyxy boxes); survivors ordered by descending score.
 * detections [batch][A][7] = (x1, y1, x2, y2, obj_conf, class_conf, class_pred), det_count[batch] rows are valid.
 * mutate_prediction != 0 also rewrites prediction[..., :4] to xyxy in place, as the reference does (boxes.py:177).     */
int64_t yb200_nms_workspace(int batch, int num_anchors);
int yb200_postprocess_nms(float* prediction, int batch, int num_anchors, int num_classes, float conf_thre, float nms_thre,
                          int mutate_prediction, void* workspace, float* detections, int32_t* det_count, void* stream);
/* Same, plus (both nullable): det_anchor [batch][A] = anchor index of every emitted row, and tie_count[batch] = number of
 * adjacent emitted rows with bit-identical scores.  Equal scores are emitted lower-anchor-first (the stable order of
 * torchvision's `nms`, i.e. the reference on CUDA and on CPU with <= 1000 candidates); the CPU `_batched_nms_vanilla`
 * path ends with an UNSTABLE torch.sort whose permutation of ties a caller can re-apply from det_anchor when
 * tie_count > 0 (yolov7_d2_b200.modeling.postprocess(tie_order="torch_cpu_sort")).                                      */
int yb200_postprocess_nms_indexed(float* prediction, int batch, int num_anchors, int num_classes, float conf_thre,
                                  float nms_thre, int mutate_prediction, void* workspace, float* detections,
                                  int32_t* det_count, int32_t* det_anchor, int32_t* tie_count, void* stream);

/* ---- box regression losses ----------------------------------------------------------------------- */
/* loss[i] and d loss[i] / d pred[i] for n matched (prediction, target) pairs of (cx, cy, w, h) boxes (device fp32 [n][4]).
 * mode 0: IOUloss "iou" = 1 - iou^2 (boxes.py:125-151, the YOLOX-s loss, yolox_head.py:134); 1: IOUloss "giou" (boxes.py:152-161);
 * 2 / 3 / 4: IOUlossV6 giou / diou / ciou with eps 1e-7 (boxes.py:666-752, YOLOv6 head).  dloss_dpred may be NULL.            */
int yb200_iou_loss(const float* pred_cxcywh, const float* target_cxcywh, int n, int mode, float* loss, float* dloss_dpred,
                   void* stream);

/* ---- parameter update on the flat fp32 buffers -------------------------------------------------------
 * Per-parameter hyper-parameters come from a segment table in device memory: seg_begin[nseg] ascending element offsets
 * (seg_begin[0] == 0), seg_wd[nseg] weight decay, seg_lr_mult[nseg] learning-rate multiplier (NULL = 1): the param groups that
 * yolov7/optimizer/build.py:77-170 builds (norm / bias / embedding decay, bias_lr_factor, lr_multipliers_overwrite).
 * grad_scale multiplies every gradient first (1/world_size of the DDP mean); total_norm (device scalar from yb200_grad_norm, may be
 * NULL) with max_norm > 0 applies clip_grad_norm_ over the whole model (FullModelGradientClippingOptimizer, build.py:206-223).       */
int64_t yb200_grad_norm_workspace(void);
/* out_norm[0] = || grad * grad_scale ||_2, deterministic two-stage reduction (fp64 accumulation) */
int yb200_grad_norm(const float* grad, int64_t n, float grad_scale, void* workspace, float* out_norm, void* stream);
/* torch.optim.SGD.step (build.py:234-245): d = g + wd*p; buf = first_step ? d : momentum*buf + (1-dampening)*d;
 * d = nesterov ? d + momentum*buf : buf; p -= lr*d.  momentum_buf may be NULL when momentum == 0.                                   */
int yb200_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t n, const int64_t* seg_begin, const float* seg_wd,
                   const float* seg_lr_mult, int nseg, float lr, float momentum, float dampening, int nesterov, int first_step,
                   float grad_scale, const float* total_norm, float max_norm, void* stream);
/* torch.optim.AdamW.step (build.py:248-256): p *= 1 - lr*wd; m, v moments; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps); step t >= 1 */
int yb200_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, const int64_t* seg_begin,
                     const float* seg_wd, const float* seg_lr_mult, int nseg, float lr, float beta1, float beta2, float eps, int step,
                     float grad_scale, const float* total_norm, float max_norm, void* stream);

/* ---- ConvNeXt block (yolov7/modeling/backbone/convnext.py) ------------------------------------------------
 * The two Linear layers of the block and the 2x2 / 4x4 strided convolutions run on the implicit-GEMM kernels above with these
 * epilogues; everything else is in csrc/convnext.cu.  Linear weights [out][in] are packed with yb200_pack_conv_weight(ksize 1).      */
/* as yb200_pack_conv_weight with every output row scaled by cout_scale[co] first (layer scale gamma folded into pwconv2, convnext.py:54-56);
 * scaled_bias[co] = cout_scale[co] * bias[co] (both NULL or both given): the shift of the matching yb200_conv2d_affine_fwd call              */
int yb200_pack_conv_weight_scaled(const float* w_oihw, const float* cout_scale, const float* bias, int cout, int cin, int ksize, int cout_pad,
                                  int cin_pad, void* w_fwd, void* w_dgrad, float* scaled_bias, void* stream);
/* out = bf16(conv(x) * scale[c] + shift[c] [+ residual]); scale / shift may be NULL (1 / 0).  ksize/stride: 1/1, 3/1, 3/2 or 2/2 (no padding:
 * the downsample convolutions, convnext.py:86-91).  nn.Conv2d / nn.Linear with bias: shift = bias.                                      */
int yb200_conv2d_affine_fwd(const yb200_act* x, const void* w_fwd, const float* scale, const float* shift, const yb200_act* residual,
                            const yb200_act* out, int ksize, int stride, void* stream);
/* pwconv1 + GELU (convnext.py:52-53): u = bf16(x W^T + bias) -> u_out (may be NULL), h = bf16(GELU(u)) -> h_out (exact erf form)       */
int yb200_linear_gelu_fwd(const yb200_act* x, const void* w_fwd, const float* bias, const yb200_act* u_out, const yb200_act* h_out, void* stream);
/* du = bf16((dz W) * GELU'(u)) and, when bias_grad_sum != NULL, bias_grad_sum[c] += sum over pixels of du[.., c] (fp64, caller zeroes it):
 * the data gradient of pwconv2 fused with the GELU backward and the bias gradient of pwconv1.                                            */
int yb200_linear_dgrad_gelu(const yb200_act* dz, const void* w_dgrad, const yb200_act* u, const yb200_act* du, double* bias_grad_sum,
                            void* stream);
/* depthwise 7x7, stride 1, zero padding 3 (convnext.py:41,49): out = dwconv(x; w) [+ bias] [+ addend].  w_c49: fp32 [C][7][7] (the
 * nn.Conv2d(groups=C) weight [C][1][7][7]).  flip != 0 correlates with the flipped kernel = the data gradient (addend: the residual branch). */
int yb200_dwconv7(const yb200_act* x, const float* w_c49, const float* bias, const yb200_act* addend, const yb200_act* out, int flip,
                  void* stream);
int64_t yb200_dwconv7_wgrad_workspace(const yb200_act* x);
/* grad_w_c49[c][ky][kx] = sum dy[p][c] x[p + (ky-3, kx-3)][c]; grad_bias[c] = sum dy[p][c] (may be NULL); fixed summation order          */
int yb200_dwconv7_wgrad(const yb200_act* x, const yb200_act* dy, float* grad_w_c49, float* grad_bias, int accumulate, void* workspace,
                        void* stream);
/* LayerNorm over the channel dimension of every pixel (convnext.py:196-206, both data formats; biased variance, eps inside the sqrt).
 * stats_mean_rstd: fp32 [pixels][2], may be NULL in forward-only use.                                                                    */
int yb200_layernorm_fwd(const yb200_act* x, const float* gamma, const float* beta, float eps, const yb200_act* y, float* stats_mean_rstd,
                        void* stream);
int64_t yb200_layernorm_bwd_workspace(const yb200_act* x);
/* dx = LayerNorm-backward(dy) [+ addend]; grad_gamma / grad_beta fp32 [C], fixed summation order                                         */
int yb200_layernorm_bwd(const yb200_act* dy, const yb200_act* x, const float* stats_mean_rstd, const float* gamma, const yb200_act* addend,
                        const yb200_act* dx, float* grad_gamma, float* grad_beta, int accumulate, void* workspace, void* stream);
int64_t yb200_colsum_workspace(const yb200_act* x);
/* out[c] = scale * sum over pixels of x[.., c] (bias gradients), fixed summation order                                                    */
int yb200_colsum(const yb200_act* x, float scale, float* out, int accumulate, void* workspace, void* stream);
/* Layer scale gamma (convnext.py:44-45,55-56) folded into pwconv2: given raw_wgrad = dOut^T h (weight gradient w.r.t. the UNscaled output
 * gradient, [C][hidden]) and gout_colsum[c] = sum of dOut:  grad_w2 = gamma[c] * raw, grad_gamma[c] = <w2[c], raw[c]> + b2[c] * colsum[c],
 * grad_b2[c] = gamma[c] * colsum[c].  raw_wgrad may alias grad_w2.                                                                         */
int yb200_layer_scale_grad(const float* raw_wgrad, const float* w2, const float* b2, const float* gamma, const float* gout_colsum, int channels,
                           int hidden, float* grad_w2, float* grad_gamma, float* grad_b2, int accumulate, void* stream);
/* stem input (convnext.py:81-84): uint8 (is_f32 = 0) or fp32 NCHW image -> [N][H/4][W/4][48] bf16 patches, channel = c*16 + kh*4 + kw (the OIHW flattening of
 * the 4x4 stride-4 stem weight), so the stem is a K = 48 GEMM (yb200_conv2d_affine_fwd, ksize 1).                                          */
int yb200_patchify4(const void* images_nchw, int is_f32, int n, int h, int w, const yb200_act* out, void* stream);
/* dst[i] = (accumulate ? dst[i] : 0) + (float)src[i]; src[i] = 0 when zero_src (fp64 accumulators of the GEMM epilogues -> fp32 gradients) */
int yb200_f64_to_f32(double* src, int n, float* dst, int accumulate, int zero_src, void* stream);

/* ---- DETR transformer attention (yolov7/modeling/backbone/detr_backbone.py:140,157-161,200-236) -------------
 * The core of nn.MultiheadAttention between in_proj and out_proj, head dimension 32:
 *   out[b, i, h, :] = softmax_j(scale * <q[b,i,h,:], k[b,j,h,:]> + (key_padding_mask[b,j] ? -inf : 0)) . v[b,j,h,:]
 * q / out: views [B][1][Lq][heads*32], k / v: [B][1][Lk][heads*32] (token-major, head h = channels [32h, 32h+32) of the view; q, k, v may be
 * slices of one packed in_proj output).  key_padding_mask: uint8 [B][Lk], 1 = ignore, or NULL.  scale: head_dim^-0.5 in the reference.
 * lse (fp32 [B][heads][Lq], may be NULL) receives log sum_j exp(scaled masked score), the statistic a backward pass needs.
 * A query whose keys are all masked yields zeros (torch yields NaN).  in_proj / out_proj / FFN are yb200_conv2d_affine_fwd calls.          */
int yb200_attention_fwd(const yb200_act* q, const yb200_act* k, const yb200_act* v, const uint8_t* key_padding_mask, float scale,
                        const yb200_act* out, float* lse, void* stream);
/* Linear + ReLU of the transformer FFN (detr_backbone.py:167, 239): h = bf16(max(x W^T + bias, 0))                                        */
int yb200_linear_relu_fwd(const yb200_act* x, const void* w_fwd, const float* bias, const yb200_act* h_out, void* stream);
/* du = bf16(h > 0 ? dz W : 0) and optionally bias_grad_sum[c] += column sums of du (fp64): data gradient of linear2 fused with ReLU backward */
int yb200_linear_dgrad_relu(const yb200_act* dz, const void* w_dgrad, const yb200_act* h, const yb200_act* du, double* bias_grad_sum, void* stream);
/* out = a + b (bf16 views of equal shape): tensor + positional embedding (detr_backbone.py:154-155, 218-219)                                */
int yb200_add(const yb200_act* a, const yb200_act* b, const yb200_act* out, void* stream);

/* ---- SparseInst IAM decoder forward (yolov7/modeling/transcoders/decoder_sparseinst.py:27-169) ------------------
 * 3x3 convolution + bias + ReLU (`_make_stack_3x3_convs` :18-24); ksize / stride as yb200_conv2d_affine_fwd                                 */
int yb200_conv2d_relu_fwd(const yb200_act* x, const void* w_fwd, const float* bias, const yb200_act* out, int ksize, int stride, void* stream);
/* 1x1 convolution with fp32 NCHW output [N][cout][H][W] (+ bias, may be NULL): with per-image weights = pred_kernel[b] this is
 * `torch.bmm(pred_kernel, mask_features.view(B, C, HW))` (:143-146); cout <= 128                                                            */
int yb200_conv1x1_nchw_f32(const yb200_act* x, const void* w_fwd, const float* bias, int cout, float* out_nchw, void* stream);
/* The same for a whole batch with ONE WEIGHT MATRIX PER IMAGE (w_fwd: [x->n][cout][x->c] bf16 = the batch's predicted kernels): the batched
 * `torch.bmm` of :143-146 in one launch.  h * w must be a multiple of 128 pixels arranged so that a tile stays inside one image.            */
int yb200_conv1x1_nchw_f32_batched(const yb200_act* x, const void* w_fwd, int cout, float* out_nchw, void* stream);
/* F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False) on fp32 planes [planes][h][w] -> [planes][2h][2w]: the mask logits
 * (:148-153)                                                                                                                                  */
int yb200_upsample_bilinear2x_f32(const float* in, float* out, int64_t planes, int h, int w, void* stream);
/* out = sigmoid(x): instance activation maps (:67)                                                                                           */
int yb200_sigmoid(const yb200_act* x, const yb200_act* out, void* stream);
/* inst[r][c] = raw[r][c] / max(normalizer[r], 1e-6) -> bf16 [1][1][rows][cols] view (:75-76).  raw = iam_prob^T features of one image is
 * yb200_conv2d_wgrad(x = features, dz = iam_prob, ksize 1) (the pixel contraction of :74), normalizer = yb200_colsum(iam_prob).              */
int yb200_iam_normalize(const float* raw, const float* normalizer, int rows, int cols, const yb200_act* out, void* stream);
/* Backward of yb200_attention_fwd: given out, its gradient dout and the saved lse, dq / dk / dv (bf16 views shaped like q / k / k; they may be
 * slices of one packed buffer).  P is recomputed from lse; two kernels (per key tile: dK, dV; per query tile: dQ), accumulation in TMEM,
 * no atomics.  workspace: yb200_attention_bwd_workspace(q) bytes (D = <dout, out> per query row and head).                                     */
int64_t yb200_attention_bwd_workspace(const yb200_act* q);
int yb200_attention_bwd(const yb200_act* q, const yb200_act* k, const yb200_act* v, const yb200_act* out, const yb200_act* dout,
                        const uint8_t* key_padding_mask, float scale, const float* lse, const yb200_act* dq, const yb200_act* dk,
                        const yb200_act* dv, void* workspace, void* stream);

/* Dropout of the transformer layers (detr_backbone.py:132-152, 200-214).  The masks are a counter-based hash of (seed, element index) evaluated
 * inside the kernels -- keep(element) with probability 1 - p_drop, kept values scaled by 1 / (1 - p_drop) -- so the backward calls regenerate the
 * mask of the forward from the same seed; p_drop = 0 is exactly the dropout-free kernel.  (Not torch's Philox stream: same distribution, different
 * bits; oracle/detr_oracle.py restates the hash for the parity tests.)
 *   yb200_attention_fwd_dropout / _bwd_dropout: nn.MultiheadAttention(dropout = p) -- the mask multiplies the softmax probabilities (index: image, head,
 *     query, key); lse stays the log-sum-exp of the undropped scores.
 *   yb200_dropout: out = residual + x * mask(seed) / (1 - p) * extra_scale on [B][1][L][C] bf16 views (index: logical element ((b L + l) C + c));
 *     residual may be NULL; the backward of nn.Dropout is the same call on the gradient.                                                          */
int yb200_attention_fwd_dropout(const yb200_act* q, const yb200_act* k, const yb200_act* v, const uint8_t* key_padding_mask, float scale,
                                const yb200_act* out, float* lse, float p_drop, uint32_t seed, void* stream);
int yb200_attention_bwd_dropout(const yb200_act* q, const yb200_act* k, const yb200_act* v, const yb200_act* out, const yb200_act* dout,
                                const uint8_t* key_padding_mask, float scale, const float* lse, const yb200_act* dq, const yb200_act* dk,
                                const yb200_act* dv, void* workspace, float p_drop, uint32_t seed, void* stream);
int yb200_dropout(const yb200_act* x, const yb200_act* residual, const yb200_act* out, float p_drop, uint32_t seed, float extra_scale,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YB200_H_ */
