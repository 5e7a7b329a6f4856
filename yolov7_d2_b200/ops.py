"""`torch.library` registration of the C-ABI operators ("bound via torch.library.define/impl so torch.compile / DDP see
ordinary ops").  Each op is an opaque CUDA custom op (namespace `yb200`) whose implementation is a libyb200.so call on the current
stream; fake (meta) kernels give shapes to tracing.  There are no CPU implementations: calling an op on CPU tensors raises
NotImplementedError from the dispatcher.

    torch.ops.yb200.postprocess_nms(prediction, num_classes, conf_thre, nms_thre) -> (detections [B,A,7], counts [B])
        `postprocess` (yolov7/utils/boxes.py:171-210) for the whole batch; rewrites prediction[..., :4] to corners in place
    torch.ops.yb200.conv2d_bn_silu(x, w_packed, scale, shift, residual, cout, ksize, stride) -> out
        eval-mode BaseConv (wrappers.py:60-83) [+ Bottleneck shortcut wrappers.py:119-123]; x / out NHWC bf16
    torch.ops.yb200.conv2d_bn_silu_train(x, w_packed, cout, ksize, stride) -> (z fp16 NHWC, sum fp64 [cout], sqsum fp64 [cout])
        training-mode convolution with the BatchNorm batch statistics from its epilogue (wrappers.py:67-80)
    torch.ops.yb200.pack_conv_weight(w_oihw, cin_pad) -> packed bf16 [cout, k*k, cin_pad]
    torch.ops.yb200.iou_loss(pred, target, mode) -> (loss [n], dloss/dpred [n,4])     IOUloss / IOUlossV6 (boxes.py:125-168, 666-752)
    torch.ops.yb200.iou_loss_autograd(pred, target, mode) -> loss [n]                differentiable w.r.t. pred (autograd.Function on top)
"""
import ctypes

import torch

from . import capi

_f = ctypes.c_float


def _act(t):
    return capi.act(t)


@torch.library.custom_op("yb200::postprocess_nms", mutates_args={"prediction"}, device_types="cuda")
def postprocess_nms(prediction: torch.Tensor, num_classes: int, conf_thre: float, nms_thre: float) -> tuple[torch.Tensor, torch.Tensor]:
    assert prediction.dtype == torch.float32 and prediction.dim() == 3 and prediction.is_contiguous()
    b, a, ch = prediction.shape
    if ch != 5 + num_classes:
        raise IndexError(f"prediction has {ch} channels, expected {5 + num_classes}")
    L = capi.lib()
    ws = torch.empty(L.yb200_nms_workspace(b, a), dtype=torch.uint8, device=prediction.device)
    det = torch.zeros(b, a, 7, device=prediction.device)
    cnt = torch.empty(b, dtype=torch.int32, device=prediction.device)
    capi.check(L.yb200_postprocess_nms(capi.ptr(prediction), b, a, num_classes, _f(conf_thre), _f(nms_thre), 1, capi.ptr(ws), capi.ptr(det), capi.ptr(cnt),
                                       capi.stream_ptr()), "postprocess_nms")
    return det, cnt


@postprocess_nms.register_fake
def _(prediction, num_classes, conf_thre, nms_thre):
    b, a, _ = prediction.shape
    return prediction.new_empty(b, a, 7), prediction.new_empty(b, dtype=torch.int32)


@torch.library.custom_op("yb200::pack_conv_weight", mutates_args=(), device_types="cuda")
def pack_conv_weight(w_oihw: torch.Tensor, cin_pad: int) -> torch.Tensor:
    cout, cin, k, _ = w_oihw.shape
    out = torch.empty(cout, k * k, cin_pad, dtype=torch.bfloat16, device=w_oihw.device)
    capi.check(capi.lib().yb200_pack_conv_weight(capi.ptr(w_oihw.contiguous().float()), cout, cin, k, cout, cin_pad, capi.ptr(out), None, capi.stream_ptr()),
               "pack_conv_weight")
    return out


@pack_conv_weight.register_fake
def _(w_oihw, cin_pad):
    cout, _, k, _ = w_oihw.shape
    return w_oihw.new_empty(cout, k * k, cin_pad, dtype=torch.bfloat16)


@torch.library.custom_op("yb200::conv2d_bn_silu", mutates_args=(), device_types="cuda")
def conv2d_bn_silu(x: torch.Tensor, w_packed: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, residual: torch.Tensor | None, cout: int,
                   ksize: int, stride: int) -> torch.Tensor:
    n, h, w, _ = x.shape
    out = torch.empty(n, h // stride, w // stride, cout, dtype=torch.bfloat16, device=x.device)
    xa, oa = _act(x), _act(out)
    ra = _act(residual) if residual is not None else None
    capi.check(capi.lib().yb200_conv2d_bn_silu_fwd(ctypes.byref(xa), capi.ptr(w_packed), capi.ptr(scale), capi.ptr(shift),
                                                   ctypes.byref(ra) if ra is not None else None, ctypes.byref(oa), ksize, stride, capi.stream_ptr()),
               "conv2d_bn_silu")
    return out


@conv2d_bn_silu.register_fake
def _(x, w_packed, scale, shift, residual, cout, ksize, stride):
    n, h, w, _ = x.shape
    return x.new_empty(n, h // stride, w // stride, cout)


@torch.library.custom_op("yb200::conv2d_bn_silu_train", mutates_args=(), device_types="cuda")
def conv2d_bn_silu_train(x: torch.Tensor, w_packed: torch.Tensor, cout: int, ksize: int, stride: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    n, h, w, _ = x.shape
    z = torch.empty(n, h // stride, w // stride, cout, dtype=torch.float16, device=x.device)
    s1 = torch.zeros(cout, dtype=torch.float64, device=x.device)
    s2 = torch.zeros(cout, dtype=torch.float64, device=x.device)
    xa, za = _act(x), _act(z)
    capi.check(capi.lib().yb200_conv2d_fwd(ctypes.byref(xa), capi.ptr(w_packed), ctypes.byref(za), ksize, stride, capi.ptr(s1), capi.ptr(s2),
                                           capi.stream_ptr()), "conv2d_fwd")
    return z, s1, s2


@conv2d_bn_silu_train.register_fake
def _(x, w_packed, cout, ksize, stride):
    n, h, w, _ = x.shape
    return (x.new_empty(n, h // stride, w // stride, cout, dtype=torch.float16), x.new_empty(cout, dtype=torch.float64),
            x.new_empty(cout, dtype=torch.float64))


@torch.library.custom_op("yb200::iou_loss", mutates_args=(), device_types="cuda")
def iou_loss(pred: torch.Tensor, target: torch.Tensor, mode: int) -> tuple[torch.Tensor, torch.Tensor]:
    if pred.shape[-1] != 4 or target.shape != pred.shape:
        raise IndexError("iou_loss expects [n, 4] boxes")
    p, t = pred.contiguous().float(), target.contiguous().float()
    n = p.shape[0]
    loss = torch.empty(n, device=p.device)
    grad = torch.empty(n, 4, device=p.device)
    if n:
        capi.check(capi.lib().yb200_iou_loss(capi.ptr(p), capi.ptr(t), n, mode, capi.ptr(loss), capi.ptr(grad), capi.stream_ptr()), "iou_loss")
    return loss, grad


@iou_loss.register_fake
def _(pred, target, mode):
    return pred.new_empty(pred.shape[0]), pred.new_empty(pred.shape[0], 4)


def _iou_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])


def _iou_backward(ctx, g_loss, g_grad):
    (dpred,) = ctx.saved_tensors
    return g_loss[:, None] * dpred, None, None


iou_loss.register_autograd(_iou_backward, setup_context=_iou_setup)


def iou_loss_autograd(pred, target, mode=0):
    """loss [n], differentiable w.r.t. pred: `IOUloss(reduction="none")` (mode 0 iou^2, 1 giou) / `IOUlossV6` (2 giou, 3 diou, 4 ciou)"""
    return torch.ops.yb200.iou_loss(pred, target, mode)[0]
