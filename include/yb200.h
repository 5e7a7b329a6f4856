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
 * ksize in {1,3}, stride in {1,2} (stride 2 only with ksize 3).  z is the bf16 pre-BatchNorm output.
 * If stat_sum/stat_sqsum are non-NULL the per-channel sum and sum of squares of the *stored* z are
 * accumulated into them (fp64, must be zeroed by the caller) -- the batch statistics nn.BatchNorm2d
 * (wrappers.py:76) computes in training mode.                                                          */
int yb200_conv2d_fwd(const yb200_act* x, const void* w_fwd, const yb200_act* z, int ksize, int stride,
                     double* stat_sum, double* stat_sqsum, void* stream);

/* out[n, a_off + y*w + x, c_off + c] = conv1x1(x, w)[n,y,x,c] + bias[c] in fp32 -- the prediction convs
 * yolox_head.py:103-129 fused with the cat/flatten/permute of yolox_head.py:175,238-244.
 * out is [n][a_total][c_total] fp32.                                                                    */
int yb200_conv1x1_bias_f32(const yb200_act* x, const void* w_fwd, const float* bias, int cout, float* out,
                           int a_total, int a_off, int c_total, int c_off, void* stream);

/* dx = conv_transpose(dz, w) [+ addend] -- autograd of the convolution w.r.t. its input.  dx/addend have the
 * input's shape, dz the output's.  addend may be NULL.                                                   */
int yb200_conv2d_dgrad(const yb200_act* dz, const void* w_dgrad, const yb200_act* dx, const yb200_act* addend,
                       int ksize, int stride, void* stream);

/* grad_oihw (+)= d loss / d weight, fp32 [cout][cin_real][k][k] -- autograd of the convolution w.r.t. its weight.
 * workspace: at least yb200_conv2d_wgrad_workspace() bytes.                                              */
int64_t yb200_conv2d_wgrad_workspace(const yb200_act* x, const yb200_act* dz, int ksize, int stride);
int yb200_conv2d_wgrad(const yb200_act* x, const yb200_act* dz, int ksize, int stride, int cin_real,
                       float* grad_oihw, int accumulate, void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YB200_H_ */
