"""A/B timing of the GEMM epilogues on ConvNeXt shapes (CUDA events): plain bf16 store vs bias+GELU with one or two outputs vs GELU backward."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolov7_d2_b200 import capi

dev = torch.device("cuda:0")
L = capi.lib()

def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for (n, h, w, c) in [(32, 160, 160, 96), (32, 80, 80, 192), (32, 40, 40, 384), (32, 20, 20, 768)]:
    hid = 4 * c
    x = torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)
    w1 = torch.randn(hid, c, 1, 1, device=dev) / c ** 0.5
    b1 = torch.randn(hid, device=dev)
    wf = torch.empty(hid, 1, c, dtype=torch.bfloat16, device=dev)
    wd = torch.empty(c, 1, hid, dtype=torch.bfloat16, device=dev)
    capi.check(L.yb200_pack_conv_weight(capi.ptr(w1), hid, c, 1, hid, c, capi.ptr(wf), capi.ptr(wd), capi.stream_ptr()), "pack")
    u = torch.empty(n, h, w, hid, dtype=torch.bfloat16, device=dev)
    hh = torch.empty_like(u)
    du = torch.empty_like(u)
    g = torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)
    w2 = torch.randn(c, hid, 1, 1, device=dev) / hid ** 0.5
    w2f = torch.empty(c, 1, hid, dtype=torch.bfloat16, device=dev)
    w2d = torch.empty(hid, 1, c, dtype=torch.bfloat16, device=dev)
    capi.check(L.yb200_pack_conv_weight(capi.ptr(w2), c, hid, 1, c, hid, capi.ptr(w2f), capi.ptr(w2d), capi.stream_ptr()), "pack")
    acc = torch.zeros(hid, dtype=torch.float64, device=dev)
    xa, ua, ha, ga, dua = capi.act(x), capi.act(u), capi.act(hh), capi.act(g), capi.act(du)
    sp = capi.stream_ptr()
    R = ctypes.byref
    t_plain = timeit(lambda: capi.check(L.yb200_conv2d_affine_fwd(R(xa), capi.ptr(wf), None, None, None, R(ha), 1, 1, sp), "a"))
    t_bias = timeit(lambda: capi.check(L.yb200_conv2d_affine_fwd(R(xa), capi.ptr(wf), None, capi.ptr(b1), None, R(ha), 1, 1, sp), "b"))
    t_g1 = timeit(lambda: capi.check(L.yb200_linear_gelu_fwd(R(xa), capi.ptr(wf), capi.ptr(b1), None, R(ha), sp), "c"))
    t_g2 = timeit(lambda: capi.check(L.yb200_linear_gelu_fwd(R(xa), capi.ptr(wf), capi.ptr(b1), R(ua), R(ha), sp), "d"))
    t_dg = timeit(lambda: capi.check(L.yb200_conv2d_dgrad(R(ga), capi.ptr(w2d), R(dua), None, 1, 1, sp), "e"))
    t_dgg = timeit(lambda: capi.check(L.yb200_linear_dgrad_gelu(R(ga), capi.ptr(w2d), R(ua), R(dua), capi.ptr(acc), sp), "f"))
    elems = n * h * w * hid
    print("C=%4d M=%7d: plain %7.1f us (%.0f Gelem/s) | +bias %7.1f | gelu 1 out %7.1f | gelu 2 out %7.1f | dgrad plain %7.1f | dgrad gelu' %7.1f | HBM floor (2 out) %.0f us"
          % (c, n * h * w, t_plain, elems / t_plain / 1e3, t_bias, t_g1, t_g2, t_dg, t_dgg, (elems * 4 + n * h * w * c * 2) / 6.5e6))
