"""GPU parity of the ConvNeXt block kernels through the C ABI.

Reference for every op: torch fp32 on the same bf16-rounded inputs (TF32 off) -- the arithmetic yolov7/modeling/backbone/convnext.py
performs (nn.Conv2d groups=C, F.layer_norm, nn.Linear, nn.GELU, layer scale) and its autograd.
Tolerances (written at each check): tensors stored in bf16 carry rel 2^-8 rounding => 2^-7 of the tensor's max; fp32 parameter
gradients accumulate in fp32 in both paths => 2e-3 of the tensor's max (different summation orders over up to 1e5 pixels).
"""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF = 2.0 ** -7


def _close(got, ref, tol, what):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert torch.isfinite(got).all(), f"{what}: non-finite values"
    assert err <= tol * max(scale, 1e-6), f"{what}: max err {err:.4e} > {tol:.1e} * {scale:.4e}"


def _bf(t):
    return t.to(torch.bfloat16)


def _nhwc(n, h, w, c, dev, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return _bf((torch.randn(n, h, w, c, generator=g) * scale).to(dev))


def _nchw(t):
    return t.float().permute(0, 3, 1, 2)


def _ws(nbytes, dev):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)


DW_SHAPES = [(2, 20, 44, 96), (1, 8, 32, 32), (3, 13, 7, 64), (1, 40, 40, 192)]


@pytest.mark.parametrize("shape", DW_SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("flip", [0, 1])
def test_dwconv7(cuda, shape, flip):
    from yolov7_d2_b200 import capi

    n, h, w, c = shape
    x = _nhwc(n, h, w, c, cuda, 1)
    add = _nhwc(n, h, w, c, cuda, 2)
    g = torch.Generator().manual_seed(3)
    wt = (torch.randn(c, 1, 7, 7, generator=g) * 0.15).to(cuda)
    bias = (torch.randn(c, generator=g) * 0.2).to(cuda)
    out = torch.full((n, h, w, c), float("nan"), dtype=torch.bfloat16, device=cuda)
    xa, aa, oa = capi.act(x), capi.act(add), capi.act(out)
    capi.check(capi.lib().yb200_dwconv7(ctypes.byref(xa), capi.ptr(wt), None if flip else capi.ptr(bias), ctypes.byref(aa) if flip else None,
                                        ctypes.byref(oa), flip, capi.stream_ptr()), "dwconv7")
    if flip:  # data gradient of the forward op w.r.t. its input, plus the residual branch
        xin = _nchw(x).requires_grad_(True)
        F.conv2d(xin, wt, None, padding=3, groups=c).backward(_nchw(x))  # any upstream gradient: reuse x
        # dx = conv_transpose(dy, w): feed dy = x
        ref = F.conv_transpose2d(_nchw(x), wt, None, padding=3, groups=c) + _nchw(add)
        _close(_nchw(out), ref, BF, "dwconv7 data gradient")
        _close(xin.grad, F.conv_transpose2d(_nchw(x), wt, None, padding=3, groups=c), 1e-5, "autograd cross-check")
    else:
        ref = F.conv2d(_nchw(x), wt, bias, padding=3, groups=c)
        _close(_nchw(out), ref, BF, "dwconv7 forward")


@pytest.mark.parametrize("shape", DW_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_dwconv7_wgrad(cuda, shape):
    from yolov7_d2_b200 import capi

    n, h, w, c = shape
    x = _nhwc(n, h, w, c, cuda, 4)
    dy = _nhwc(n, h, w, c, cuda, 5)
    L = capi.lib()
    xa, da = capi.act(x), capi.act(dy)
    ws = _ws(L.yb200_dwconv7_wgrad_workspace(ctypes.byref(xa)), cuda)
    gw = torch.full((c, 1, 7, 7), float("nan"), device=cuda)
    gb = torch.full((c,), float("nan"), device=cuda)
    capi.check(L.yb200_dwconv7_wgrad(ctypes.byref(xa), ctypes.byref(da), capi.ptr(gw), capi.ptr(gb), 0, capi.ptr(ws), capi.stream_ptr()), "dwconv7_wgrad")
    wt = torch.zeros(c, 1, 7, 7, device=cuda, requires_grad=True)
    b = torch.zeros(c, device=cuda, requires_grad=True)
    F.conv2d(_nchw(x), wt, b, padding=3, groups=c).backward(_nchw(dy))
    _close(gw, wt.grad, 2e-3, "dwconv7 weight gradient")
    _close(gb, b.grad, 2e-3, "dwconv7 bias gradient")
    # accumulate adds, and the result is bit-reproducible
    gw2, gb2 = gw.clone(), gb.clone()
    capi.check(L.yb200_dwconv7_wgrad(ctypes.byref(xa), ctypes.byref(da), capi.ptr(gw2), capi.ptr(gb2), 1, capi.ptr(ws), capi.stream_ptr()), "dwconv7_wgrad acc")
    assert torch.equal(gw2, 2 * gw) and torch.equal(gb2, 2 * gb)


LN_SHAPES = [(2, 9, 11, 96), (1, 5, 7, 24), (2, 6, 6, 192), (1, 7, 9, 384), (1, 4, 5, 768), (3, 3, 3, 1024)]


@pytest.mark.parametrize("shape", LN_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_layernorm_fwd_bwd(cuda, shape):
    from yolov7_d2_b200 import capi

    n, h, w, c = shape
    L = capi.lib()
    x = _nhwc(n, h, w, c, cuda, 6, scale=2.0) + 0.5
    x = _bf(x)
    dy = _nhwc(n, h, w, c, cuda, 7)
    add = _nhwc(n, h, w, c, cuda, 8)
    g = torch.Generator().manual_seed(9)
    gamma = (torch.rand(c, generator=g) + 0.5).to(cuda)
    beta = (torch.rand(c, generator=g) - 0.5).to(cuda)
    y = torch.full_like(x, float("nan"))
    stats = torch.full((n * h * w, 2), float("nan"), device=cuda)
    xa, ya, da, aa = capi.act(x), capi.act(y), capi.act(dy), capi.act(add)
    capi.check(L.yb200_layernorm_fwd(ctypes.byref(xa), capi.ptr(gamma), capi.ptr(beta), ctypes.c_float(1e-6), ctypes.byref(ya), capi.ptr(stats),
                                     capi.stream_ptr()), "ln fwd")
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (c,), gr, br, 1e-6)
    _close(y, ref, BF, "LayerNorm forward")
    _close(stats[:, 0], x.float().mean(-1).reshape(-1), 1e-5, "mean")
    _close(stats[:, 1], (x.float().var(-1, unbiased=False) + 1e-6).rsqrt().reshape(-1), 1e-4, "rstd")
    ref.backward(dy.float())
    dx = torch.full_like(x, float("nan"))
    dxa = capi.act(dx)
    gg = torch.full((c,), float("nan"), device=cuda)
    gb = torch.full((c,), float("nan"), device=cuda)
    ws = _ws(L.yb200_layernorm_bwd_workspace(ctypes.byref(xa)), cuda)
    capi.check(L.yb200_layernorm_bwd(ctypes.byref(da), ctypes.byref(xa), capi.ptr(stats), capi.ptr(gamma), ctypes.byref(aa), ctypes.byref(dxa), capi.ptr(gg),
                                     capi.ptr(gb), 0, capi.ptr(ws), capi.stream_ptr()), "ln bwd")
    _close(dx, xr.grad + add.float(), BF, "LayerNorm input gradient (+addend)")
    _close(gg, gr.grad, 2e-3, "LayerNorm weight gradient")
    _close(gb, br.grad, 2e-3, "LayerNorm bias gradient")
    # in place on the addend buffer (how the engine merges two gradient branches)
    add2 = add.clone()
    a2 = capi.act(add2)
    capi.check(L.yb200_layernorm_bwd(ctypes.byref(da), ctypes.byref(xa), capi.ptr(stats), capi.ptr(gamma), ctypes.byref(a2), ctypes.byref(a2), capi.ptr(gg),
                                     capi.ptr(gb), 0, capi.ptr(ws), capi.stream_ptr()), "ln bwd in place")
    assert torch.equal(add2, dx)


def test_colsum_and_f64(cuda):
    from yolov7_d2_b200 import capi

    L = capi.lib()
    for c in (96, 80, 768, 8):
        x = _nhwc(3, 17, 9, c, cuda, 10)
        xa = capi.act(x)
        ws = _ws(L.yb200_colsum_workspace(ctypes.byref(xa)), cuda)
        out = torch.full((c,), float("nan"), device=cuda)
        capi.check(L.yb200_colsum(ctypes.byref(xa), ctypes.c_float(0.5), capi.ptr(out), 0, capi.ptr(ws), capi.stream_ptr()), "colsum")
        _close(out, 0.5 * x.float().sum((0, 1, 2)), 1e-5, f"colsum c={c}")
    src = torch.randn(1000, dtype=torch.float64, device=cuda)
    keep = src.clone()
    dst = torch.ones(1000, device=cuda)
    capi.check(L.yb200_f64_to_f32(capi.ptr(src), 1000, capi.ptr(dst), 1, 1, capi.stream_ptr()), "f64_to_f32")
    assert torch.equal(dst, 1 + keep.float()) and not src.any()


def _pack(capi, wt, scale=None, bias=None):
    cout, cin, k, _ = wt.shape
    dev = wt.device
    wf = torch.empty(cout, k * k, cin, dtype=torch.bfloat16, device=dev)
    wd = torch.empty(cin, k * k, cout, dtype=torch.bfloat16, device=dev)
    L = capi.lib()
    if scale is None:
        capi.check(L.yb200_pack_conv_weight(capi.ptr(wt), cout, cin, k, cout, cin, capi.ptr(wf), capi.ptr(wd), capi.stream_ptr()), "pack")
        return wf, wd
    sb = torch.empty(cout, device=dev) if bias is not None else None
    capi.check(L.yb200_pack_conv_weight_scaled(capi.ptr(wt), capi.ptr(scale), capi.ptr(bias), cout, cin, k, cout, cin, capi.ptr(wf), capi.ptr(wd), capi.ptr(sb),
                                               capi.stream_ptr()), "pack scaled")
    return wf, wd, sb


MLP_SHAPES = [(2, 12, 20, 32), (1, 16, 16, 96), (2, 9, 7, 192), (1, 10, 10, 384), (1, 5, 5, 768)]


@pytest.mark.parametrize("shape", MLP_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_linear_gelu_fwd_and_bwd(cuda, shape):
    """pwconv1 + GELU (convnext.py:52-53) and the fused backward of GELU behind pwconv2's data gradient"""
    from yolov7_d2_b200 import capi

    n, h, w, c = shape
    hid = 4 * c
    L = capi.lib()
    g = torch.Generator().manual_seed(11)
    x = _nhwc(n, h, w, c, cuda, 12)
    w1 = _bf((torch.randn(hid, c, 1, 1, generator=g) / c ** 0.5).to(cuda)).float()
    b1 = (torch.rand(hid, generator=g) - 0.5).to(cuda)
    wf, wd = _pack(capi, w1)
    u = torch.full((n, h, w, hid), float("nan"), dtype=torch.bfloat16, device=cuda)
    hh = torch.full_like(u, float("nan"))
    xa, ua, ha = capi.act(x), capi.act(u), capi.act(hh)
    capi.check(L.yb200_linear_gelu_fwd(ctypes.byref(xa), capi.ptr(wf), capi.ptr(b1), ctypes.byref(ua), ctypes.byref(ha), capi.stream_ptr()), "linear_gelu_fwd")
    uref = F.linear(x.float(), w1.view(hid, c), b1)
    _close(u, uref, BF, "pre-activation u")
    _close(hh, F.gelu(u.float()), BF, "GELU(u) of the stored u")
    _close(hh, F.gelu(uref), 2 * BF, "GELU(x W^T + b)")
    # backward: du = (dz W2') * GELU'(u), bias-gradient sums
    dz = _nhwc(n, h, w, c, cuda, 13)
    w2 = _bf((torch.randn(c, hid, 1, 1, generator=g) / hid ** 0.5).to(cuda)).float()
    _, w2d = _pack(capi, w2)
    du = torch.full_like(u, float("nan"))
    acc = torch.zeros(hid, dtype=torch.float64, device=cuda)
    dza, dua = capi.act(dz), capi.act(du)
    capi.check(L.yb200_linear_dgrad_gelu(ctypes.byref(dza), capi.ptr(w2d), ctypes.byref(ua), ctypes.byref(dua), capi.ptr(acc), capi.stream_ptr()), "dgrad_gelu")
    ur = u.float().requires_grad_(True)
    F.gelu(ur).backward(F.linear(dz.float(), w2.view(c, hid).t()))
    _close(du, ur.grad, BF, "du")
    _close(acc.float(), du.float().sum((0, 1, 2)), 1e-5, "bias-gradient sums equal the column sums of the stored du")


@pytest.mark.parametrize("shape", MLP_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_affine_linear_with_layer_scale_and_residual(cuda, shape):
    """pwconv2 + gamma + residual (convnext.py:54-59) with gamma folded into the packed weight, and its parameter gradients"""
    from yolov7_d2_b200 import capi

    n, h, w, c = shape
    hid = 4 * c
    L = capi.lib()
    g = torch.Generator().manual_seed(14)
    hact = _nhwc(n, h, w, hid, cuda, 15)
    res = _nhwc(n, h, w, c, cuda, 16)
    w2 = (torch.randn(c, hid, 1, 1, generator=g) / hid ** 0.5).to(cuda)
    b2 = (torch.rand(c, generator=g) - 0.5).to(cuda)
    gamma = (torch.rand(c, generator=g) * 0.5 + 0.05).to(cuda)
    wf, wd, sb = _pack(capi, w2, gamma, b2)
    out = torch.full_like(res, float("nan"))
    ha, ra, oa = capi.act(hact), capi.act(res), capi.act(out)
    capi.check(L.yb200_conv2d_affine_fwd(ctypes.byref(ha), capi.ptr(wf), None, capi.ptr(sb), ctypes.byref(ra), ctypes.byref(oa), 1, 1, capi.stream_ptr()), "affine")
    w2r, b2r, gr = w2.clone().requires_grad_(True), b2.clone().requires_grad_(True), gamma.clone().requires_grad_(True)
    hr = hact.float().requires_grad_(True)
    ref = res.float() + gr * F.linear(hr, w2r.view(c, hid), b2r)
    _close(out, ref, 1.5 * BF, "x + gamma * (h W2^T + b2)")  # gamma*W2 is rounded to bf16 once more than in the reference
    gout = _nhwc(n, h, w, c, cuda, 17)
    ref.backward(gout.float())
    # data gradient through the scaled weight
    dh = torch.full_like(hact, float("nan"))
    ga, dha = capi.act(gout), capi.act(dh)
    capi.check(L.yb200_conv2d_dgrad(ctypes.byref(ga), capi.ptr(wd), ctypes.byref(dha), None, 1, 1, capi.stream_ptr()), "dgrad")
    _close(dh, hr.grad, 1.5 * BF, "dh")
    # parameter gradients: raw weight gradient of the unscaled output gradient, then the layer-scale kernel
    ws = _ws(L.yb200_conv2d_wgrad_workspace(ctypes.byref(ha), ctypes.byref(ga), 1, 1), cuda)
    raw = torch.full((c, hid), float("nan"), device=cuda)
    capi.check(L.yb200_conv2d_wgrad(ctypes.byref(ha), ctypes.byref(ga), 1, 1, hid, capi.ptr(raw), 0, capi.ptr(ws), ctypes.c_int64(ws.numel()), capi.stream_ptr()), "wgrad")
    _close(raw, torch.einsum("nhwc,nhwk->ck", gout.float(), hact.float()), 2e-3, "raw weight gradient")
    cs = torch.empty(c, device=cuda)
    ws2 = _ws(L.yb200_colsum_workspace(ctypes.byref(ga)), cuda)
    capi.check(L.yb200_colsum(ctypes.byref(ga), ctypes.c_float(1.0), capi.ptr(cs), 0, capi.ptr(ws2), capi.stream_ptr()), "colsum")
    gw2 = torch.empty(c, hid, device=cuda)
    gg, gb2 = torch.empty(c, device=cuda), torch.empty(c, device=cuda)
    capi.check(L.yb200_layer_scale_grad(capi.ptr(raw), capi.ptr(w2), capi.ptr(b2), capi.ptr(gamma), capi.ptr(cs), c, hid, capi.ptr(gw2), capi.ptr(gg), capi.ptr(gb2), 0,
                                        capi.stream_ptr()), "layer_scale_grad")
    _close(gw2, w2r.grad.view(c, hid), 2e-3, "pwconv2 weight gradient")
    _close(gb2, b2r.grad, 2e-3, "pwconv2 bias gradient")
    _close(gg, gr.grad, 2e-3, "gamma gradient")
    # in place (raw aliases the gradient) gives the same values
    capi.check(L.yb200_layer_scale_grad(capi.ptr(raw), capi.ptr(w2), capi.ptr(b2), capi.ptr(gamma), capi.ptr(cs), c, hid, capi.ptr(raw), capi.ptr(gg), capi.ptr(gb2), 0,
                                        capi.stream_ptr()), "layer_scale_grad in place")
    assert torch.equal(raw, gw2)


DS_SHAPES = [(2, 16, 24, 32, 64), (1, 40, 40, 96, 192), (2, 10, 14, 192, 384), (1, 8, 8, 384, 768)]


@pytest.mark.parametrize("shape", DS_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_downsample_conv2x2_fwd_dgrad_wgrad(cuda, shape):
    """nn.Conv2d(kernel_size=2, stride=2) with bias (convnext.py:89): forward, data gradient (4 parity classes), weight gradient"""
    from yolov7_d2_b200 import capi

    n, h, w, cin, cout = shape
    L = capi.lib()
    g = torch.Generator().manual_seed(18)
    x = _nhwc(n, h, w, cin, cuda, 19)
    wt = _bf((torch.randn(cout, cin, 2, 2, generator=g) / (4 * cin) ** 0.5).to(cuda)).float()
    bias = (torch.rand(cout, generator=g) - 0.5).to(cuda)
    wf, wd = _pack(capi, wt)
    out = torch.full((n, h // 2, w // 2, cout), float("nan"), dtype=torch.bfloat16, device=cuda)
    xa, oa = capi.act(x), capi.act(out)
    capi.check(L.yb200_conv2d_affine_fwd(ctypes.byref(xa), capi.ptr(wf), None, capi.ptr(bias), None, ctypes.byref(oa), 2, 2, capi.stream_ptr()), "conv2x2")
    xr = _nchw(x).requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, bias, stride=2)
    _close(_nchw(out), ref, BF, "conv 2x2 s2 forward")
    dz = _nhwc(n, h // 2, w // 2, cout, cuda, 20)
    ref.backward(_nchw(dz))
    dx = torch.full_like(x, float("nan"))
    dza, dxa = capi.act(dz), capi.act(dx)
    capi.check(L.yb200_conv2d_dgrad(ctypes.byref(dza), capi.ptr(wd), ctypes.byref(dxa), None, 2, 2, capi.stream_ptr()), "dgrad 2x2")
    _close(_nchw(dx), xr.grad, BF, "conv 2x2 s2 data gradient")
    ws = _ws(L.yb200_conv2d_wgrad_workspace(ctypes.byref(xa), ctypes.byref(dza), 2, 2), cuda)
    gw = torch.full((cout, cin, 2, 2), float("nan"), device=cuda)
    capi.check(L.yb200_conv2d_wgrad(ctypes.byref(xa), ctypes.byref(dza), 2, 2, cin, capi.ptr(gw), 0, capi.ptr(ws), ctypes.c_int64(ws.numel()), capi.stream_ptr()), "wgrad 2x2")
    _close(gw, wr.grad, 2e-3, "conv 2x2 s2 weight gradient")


@pytest.mark.parametrize("is_f32", [0, 1])
def test_stem_patchify_gemm(cuda, is_f32):
    """nn.Conv2d(3, C, kernel_size=4, stride=4) (convnext.py:82) = patch gather + K=48 GEMM; weight gradient lands in OIHW order"""
    from yolov7_d2_b200 import capi

    L = capi.lib()
    n, h, w, cout = 2, 32, 48, 96
    g = torch.Generator().manual_seed(21)
    img = torch.randint(0, 256, (n, 3, h, w), generator=g, dtype=torch.uint8).to(cuda)
    src = img.float() if is_f32 else img
    patches = torch.full((n, h // 4, w // 4, 48), float("nan"), dtype=torch.bfloat16, device=cuda)
    pa = capi.act(patches)
    capi.check(L.yb200_patchify4(capi.ptr(src), is_f32, n, h, w, ctypes.byref(pa), capi.stream_ptr()), "patchify4")
    ref_p = F.unfold(img.float(), kernel_size=4, stride=4).view(n, 48, h // 4, w // 4).permute(0, 2, 3, 1)
    assert torch.equal(patches.float(), ref_p)  # 0..255 are exact in bf16
    wt = _bf((torch.randn(cout, 3, 4, 4, generator=g) * 0.02).to(cuda)).float()
    bias = (torch.rand(cout, generator=g) - 0.5).to(cuda)
    wf, _ = _pack(capi, wt.view(cout, 48, 1, 1))
    out = torch.full((n, h // 4, w // 4, cout), float("nan"), dtype=torch.bfloat16, device=cuda)
    oa = capi.act(out)
    capi.check(L.yb200_conv2d_affine_fwd(ctypes.byref(pa), capi.ptr(wf), None, capi.ptr(bias), None, ctypes.byref(oa), 1, 1, capi.stream_ptr()), "stem gemm")
    wr = wt.clone().requires_grad_(True)
    ref = F.conv2d(img.float(), wr, bias, stride=4)
    _close(_nchw(out), ref, BF, "stem forward")
    dz = _nhwc(n, h // 4, w // 4, cout, cuda, 22)
    ref.backward(_nchw(dz))
    dza = capi.act(dz)
    ws = _ws(L.yb200_conv2d_wgrad_workspace(ctypes.byref(pa), ctypes.byref(dza), 1, 1), cuda)
    gw = torch.full((cout, 3, 4, 4), float("nan"), device=cuda)
    capi.check(L.yb200_conv2d_wgrad(ctypes.byref(pa), ctypes.byref(dza), 1, 1, 48, capi.ptr(gw), 0, capi.ptr(ws), ctypes.c_int64(ws.numel()), capi.stream_ptr()), "stem wgrad")
    _close(gw, wr.grad, 2e-3, "stem weight gradient")


# ---------------------------------------------------------------------------------------------------------------------------------
# whole block / whole network through the engine, against the vectors the unmodified reference produced (tests/golden/convnext.npz)
# ---------------------------------------------------------------------------------------------------------------------------------
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "convnext.npz")


def _rel(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(np.asarray(b)).float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _corr(a, b):
    a, b = torch.as_tensor(a).float().cpu().flatten(), torch.as_tensor(np.asarray(b)).float().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-20)).item()


@pytest.fixture(scope="module")
def tiny(cuda):
    """the golden tiny ConvNeXt (depths 1,1,2,1; dims 16,32,48,64 are not multiples of 32 -> the engine test uses its own 32-multiple net
    checked against the oracle, which test_convnext_oracle_golden.py pins to the reference)"""
    from oracle import convnext_oracle as cnx
    from yolov7_d2_b200.convnext import ConvNeXtEngine

    depths, dims = (1, 1, 2, 1), (32, 64, 96, 128)
    sd = cnx.convnext_state_dict(7, depths=depths, dims=dims, trained_like=True)
    img = cnx.synthetic_images(2, 64, seed=11)
    eng = ConvNeXtEngine(2, 64, 64, depths, dims, 1e-6, (0, 1, 2, 3), cuda)
    eng.load_state_dict(sd)
    eng.images_u8.copy_(img.to(cuda))
    g = torch.Generator().manual_seed(13)
    # reference fp32 + emulated 16-bit storage yardstick, both from the oracle
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = img.float()
    feats = cnx.forward_features(x, sdr, depths=depths)
    gouts = [torch.randn(f.shape, generator=g).to(torch.bfloat16).float() for f in feats]
    sum((f * go).sum() for f, go in zip(feats, gouts)).backward()
    cnx.EMULATE_STORAGE = True
    try:
        sde = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        feats_e = cnx.forward_features(x, sde, depths=depths)
        sum((f * go).sum() for f, go in zip(feats_e, gouts)).backward()
    finally:
        cnx.EMULATE_STORAGE = False
    eng.pack_weights()
    outs = eng.forward_features()
    for i, go in enumerate(gouts):
        eng.stage[i].gout.t.copy_(go.permute(0, 2, 3, 1).to(cuda))
    eng.backward()
    torch.cuda.synchronize()
    return dict(eng=eng, outs=outs, feats=feats, feats_e=feats_e, sdr=sdr, sde=sde)


def test_engine_forward_features(tiny):
    """bf16 storage end to end: judged against the fp32 oracle with the error of the oracle's own 16-bit-storage emulation as yardstick"""
    for i, (o, f, fe) in enumerate(zip(tiny["outs"], tiny["feats"], tiny["feats_e"])):
        got = o.float().permute(0, 3, 1, 2)
        yard = _rel(fe.detach(), f.detach())
        err = _rel(got, f.detach())
        assert err <= max(2.0 * yard, 2.0 ** -6), f"stage {i}: rel err {err:.4f} vs emulated-storage yardstick {yard:.4f}"
        assert _corr(got, f.detach()) > 0.999


def test_engine_parameter_gradients(tiny):
    eng, sdr, sde = tiny["eng"], tiny["sdr"], tiny["sde"]
    worst = []
    for name in eng.param_names:
        ref = sdr[name].grad
        yard = _rel(sde[name].grad, ref)
        err = _rel(eng.grads[name].reshape(ref.shape), ref)
        cos = _corr(eng.grads[name], ref)
        worst.append((err / max(yard, 2.0 ** -6), name, err, yard, cos))
        assert cos > 0.995, f"{name}: cosine {cos:.5f} (rel err {err:.4f}, yardstick {yard:.4f})"
    worst.sort(reverse=True)
    ratio, name, err, yard, cos = worst[0]
    assert ratio <= 3.0, f"{name}: rel err {err:.4f} is {ratio:.1f}x the emulated-storage yardstick {yard:.4f}"


def test_engine_backward_is_reproducible_and_accumulates(tiny, cuda):
    eng = tiny["eng"]
    g1 = eng.flat_grad.clone()
    eng.backward()
    assert torch.equal(eng.flat_grad, g1), "backward is not bit-reproducible"
    eng.backward(accumulate=True)
    torch.testing.assert_close(eng.flat_grad, 2 * g1, rtol=1e-5, atol=1e-6)
    eng.backward()


def test_block_against_reference_golden(cuda):
    """one Block (dim 32) against tests/golden/convnext.npz block_* vectors produced by the reference class itself"""
    from yolov7_d2_b200 import capi

    gold = np.load(GOLD, allow_pickle=False)
    x = torch.tensor(gold["block_x"])  # [2,32,12,20]
    n, c, h, w = x.shape
    L, sp = capi.lib(), capi.stream_ptr()
    P = {k[len("block_sd/"):]: torch.tensor(gold[k]).to(cuda) for k in gold.files if k.startswith("block_sd/")}
    xb = _bf(x.permute(0, 2, 3, 1).contiguous().to(cuda))
    hid = 4 * c
    w1f, w1d = _pack(capi, P["pwconv1.weight"].view(hid, c, 1, 1))
    w2f, w2d, sb = _pack(capi, P["pwconv2.weight"].view(c, hid, 1, 1), P["gamma"], P["pwconv2.bias"])
    d, y, out = (torch.empty_like(xb) for _ in range(3))
    u = torch.empty(n, h, w, hid, dtype=torch.bfloat16, device=cuda)
    hh = torch.empty_like(u)
    stats = torch.empty(n * h * w, 2, device=cuda)
    A = {k: capi.act(v) for k, v in dict(x=xb, d=d, y=y, out=out, u=u, hh=hh).items()}
    R = {k: ctypes.byref(v) for k, v in A.items()}
    capi.check(L.yb200_dwconv7(R["x"], capi.ptr(P["dwconv.weight"]), capi.ptr(P["dwconv.bias"]), None, R["d"], 0, sp), "dw")
    capi.check(L.yb200_layernorm_fwd(R["d"], capi.ptr(P["norm.weight"]), capi.ptr(P["norm.bias"]), ctypes.c_float(1e-6), R["y"], capi.ptr(stats), sp), "ln")
    capi.check(L.yb200_linear_gelu_fwd(R["y"], capi.ptr(w1f), capi.ptr(P["pwconv1.bias"]), R["u"], R["hh"], sp), "g1")
    capi.check(L.yb200_conv2d_affine_fwd(R["hh"], capi.ptr(w2f), None, capi.ptr(sb), R["x"], R["out"], 1, 1, sp), "g2")
    ref = torch.tensor(gold["block_y"])
    got = out.float().permute(0, 3, 1, 2).cpu()
    # the block output is x + gamma * (...): compare the residual branch too, which is what the kernels compute
    assert _rel(got, ref) <= 2.0 ** -7
    br_ref = ref - x
    br_got = got - xb.float().permute(0, 3, 1, 2).cpu()
    assert _corr(br_got, br_ref) > 0.99, "residual branch of the block"
