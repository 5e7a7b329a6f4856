"""GPU parity of the HBM-bound kernels (preprocess/Focus, BN finalize/apply/backward, SPP pooling) vs torch fp32."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from oracle import yolox_oracle as orc

pytestmark = pytest.mark.gpu


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def test_preprocess_focus(cuda):
    from yolov7_d2_b200 import capi

    g = torch.Generator().manual_seed(1)
    img = torch.randint(0, 256, (3, 3, 64, 96), generator=g, dtype=torch.uint8)
    hwv = torch.tensor([[64, 96], [50, 70], [64, 40]], dtype=torch.int32)
    out = torch.empty(3, 32, 48, 16, dtype=torch.bfloat16, device=cuda)
    oa = capi.act(out)
    img_d, hwv_d = img.to(cuda), hwv.to(cuda)  # keep the device tensors alive across the asynchronous launch
    capi.check(capi.lib().yb200_preprocess_focus(capi.ptr(img_d), 3, 64, 96, capi.ptr(hwv_d), ctypes.c_float(114.0),
                                                 ctypes.byref(oa), capi.stream_ptr()), "preprocess")
    ref_in = orc.preprocess([img[i, :, :int(hwv[i, 0]), :int(hwv[i, 1])] for i in range(3)])
    ref = nhwc(orc.focus(ref_in))
    assert torch.equal(out[..., :12].float().cpu(), ref) and (out[..., 12:] == 0).all()


@pytest.mark.parametrize("c,hw,res,up", [(32, 40, False, False), (64, 20, True, False), (256, 10, False, True), (1024, 6, False, False)])
def test_bn_finalize_apply_and_backward(cuda, c, hw, res, up):
    from yolov7_d2_b200 import capi

    L = capi.lib()
    n = 4
    g = torch.Generator().manual_seed(2)
    z = (torch.randn(n, c, hw, hw, generator=g) * 1.5 + 0.3).to(torch.float16)  # pre-BN tensors are fp16
    gamma = (torch.rand(c, generator=g) + 0.5)
    beta = torch.randn(c, generator=g) * 0.2
    rm, rv = torch.randn(c, generator=g) * 0.1, torch.rand(c, generator=g) + 0.5
    resid = torch.randn(n, c, hw, hw, generator=g).to(torch.bfloat16) if res else None
    da = torch.randn(n, c, hw, hw, generator=g).to(torch.bfloat16)
    da_up = torch.randn(n, c, 2 * hw, 2 * hw, generator=g).to(torch.bfloat16) if up else None
    # ---- torch fp32 reference with autograd
    zt = z.float().requires_grad_(True)
    gt_, bt_ = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    a_ref = orc.silu(F.batch_norm(zt, rm_ref, rv_ref, gt_, bt_, True, orc.BN_MOMENTUM, orc.BN_EPS))
    y_ref = a_ref + resid.float() if res else a_ref
    loss = (y_ref * da.float()).sum()
    if up:
        loss = loss + (F.interpolate(a_ref, scale_factor=2, mode="nearest") * da_up.float()).sum()
    loss.backward()
    # ---- kernels
    zd = nhwc(z).to(cuda)
    zf = zd.double()
    ssum, ssq = zf.sum((0, 1, 2)), (zf * zf).sum((0, 1, 2))
    gd, bd, rmd, rvd = gamma.to(cuda), beta.to(cuda), rm.to(cuda), rv.to(cuda)
    nbt = torch.zeros((), dtype=torch.int64, device=cuda)
    scale, shift, mean, invstd = (torch.empty(c, device=cuda) for _ in range(4))
    capi.check(L.yb200_bn_finalize(capi.ptr(ssum), capi.ptr(ssq), c, ctypes.c_int64(n * hw * hw), capi.ptr(gd), capi.ptr(bd),
                                   ctypes.c_float(orc.BN_EPS), ctypes.c_float(orc.BN_MOMENTUM), capi.ptr(rmd), capi.ptr(rvd), capi.ptr(nbt),
                                   capi.ptr(scale), capi.ptr(shift), capi.ptr(mean), capi.ptr(invstd), capi.stream_ptr()), "bn_finalize")
    assert int(nbt) == 1 and (ssum == 0).all() and (ssq == 0).all()
    assert torch.allclose(rmd.cpu(), rm_ref, rtol=1e-5, atol=1e-6) and torch.allclose(rvd.cpu(), rv_ref, rtol=1e-5, atol=1e-6)
    out = torch.zeros(n, hw, hw, 2 * c, dtype=torch.bfloat16, device=cuda)  # write into the upper channel slice of a wider buffer
    outu = torch.zeros(n, 2 * hw, 2 * hw, c, dtype=torch.bfloat16, device=cuda) if up else None
    za, oa = capi.act(zd), capi.act(out, c, c)
    resid_d = nhwc(resid).to(cuda) if res else None
    ra = capi.act(resid_d) if res else None
    ua = capi.act(outu) if up else None
    capi.check(L.yb200_bn_apply_silu(ctypes.byref(za), capi.ptr(scale), capi.ptr(shift), ctypes.byref(ra) if res else None, ctypes.byref(oa),
                                     ctypes.byref(ua) if up else None, capi.stream_ptr()), "bn_apply_silu")
    y = out[..., c:].float().cpu()
    assert torch.allclose(y, nhwc(y_ref.detach()), rtol=2 ** -7, atol=2e-2), (y - nhwc(y_ref.detach())).abs().max()
    assert (out[..., :c] == 0).all()
    if up:
        assert torch.equal(outu.cpu(), nhwc(F.interpolate(out[..., c:].permute(0, 3, 1, 2).float(), scale_factor=2)).to(torch.bfloat16).cpu())
    # backward
    dad = nhwc(da).to(cuda)
    dz = torch.empty(zd.shape, dtype=torch.bfloat16, device=cuda)
    dgam, dbet = torch.empty(c, device=cuda), torch.empty(c, device=cuda)
    acc1, acc2 = torch.zeros(c, dtype=torch.float64, device=cuda), torch.zeros(c, dtype=torch.float64, device=cuda)
    daa, dza = capi.act(dad), capi.act(dz)
    da_up_d = nhwc(da_up).to(cuda) if up else None
    dua = capi.act(da_up_d) if up else None
    capi.check(L.yb200_bn_silu_bwd(ctypes.byref(za), ctypes.byref(daa), None, ctypes.byref(dua) if up else None, capi.ptr(scale), capi.ptr(shift),
                                   capi.ptr(mean), capi.ptr(invstd), capi.ptr(acc1), capi.ptr(acc2), ctypes.byref(dza), capi.ptr(dgam),
                                   capi.ptr(dbet), 0, capi.stream_ptr()), "bn_silu_bwd")
    ref_dz = nhwc(zt.grad)
    err = (dz.float().cpu() - ref_dz).abs().max().item()
    assert err <= 2 ** -7 * ref_dz.abs().max().item() + 1e-3, err
    assert torch.allclose(dgam.cpu(), gt_.grad, rtol=2e-3, atol=2e-2), (dgam.cpu() - gt_.grad).abs().max()
    assert torch.allclose(dbet.cpu(), bt_.grad, rtol=2e-3, atol=2e-2), (dbet.cpu() - bt_.grad).abs().max()
    assert (acc1 == 0).all() and (acc2 == 0).all()


@pytest.mark.parametrize("levels", [0, 2], ids=["random", "many_ties"])
def test_spp_pool_fwd_bwd(cuda, levels):
    """levels > 0: values quantised to multiples of 1/levels -- every window holds several equal maxima, so the gradient routing checks the
    first-maximum (row-major) rule of ATen's max_pool2d_with_indices, through the 5 -> 9 -> 13 cascade of the kernel"""
    from yolov7_d2_b200 import capi

    L = capi.lib()
    n, c, hw = 3, 64, 20
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, c, hw, hw, generator=g)
    if levels:
        x = torch.round(x * levels) / levels
        x[0, :8] = 0.0  # constant planes, +0 / -0 mixed
        x[0, :4, ::2] = -0.0
    x = x.to(torch.bfloat16)
    cat = torch.zeros(n, hw, hw, 4 * c, dtype=torch.bfloat16, device=cuda)
    cat[..., :c] = nhwc(x).to(cuda)
    arg = torch.empty(3, n, hw, hw, c, dtype=torch.uint8, device=cuda)
    views = [capi.act(cat, i * c, c) for i in range(4)]
    capi.check(L.yb200_spp_pool(ctypes.byref(views[0]), ctypes.byref(views[1]), ctypes.byref(views[2]), ctypes.byref(views[3]), capi.ptr(arg),
                                capi.stream_ptr()), "spp_pool")
    xt = x.float().requires_grad_(True)
    ref = torch.cat([xt] + [F.max_pool2d(xt, k, 1, k // 2) for k in (5, 9, 13)], 1)
    assert torch.equal(cat.float().cpu(), nhwc(ref.detach()))
    dcat = torch.randn(n, 4 * c, hw, hw, generator=g).to(torch.bfloat16)
    (ref * dcat.float()).sum().backward()
    dcd = nhwc(dcat).to(cuda)
    dviews = [capi.act(dcd, i * c, c) for i in range(4)]
    dx = torch.empty(n, hw, hw, c, dtype=torch.bfloat16, device=cuda)
    scratch = torch.empty(n * hw * hw * c, dtype=torch.float32, device=cuda)
    dxa = capi.act(dx)
    capi.check(L.yb200_spp_pool_bwd(ctypes.byref(dviews[0]), ctypes.byref(dviews[1]), ctypes.byref(dviews[2]), ctypes.byref(dviews[3]),
                                    capi.ptr(arg), capi.ptr(scratch), ctypes.byref(dxa), capi.stream_ptr()), "spp_pool_bwd")
    ref_dx = nhwc(xt.grad)
    assert torch.allclose(dx.float().cpu(), ref_dx, rtol=2 ** -7, atol=2e-2), (dx.float().cpu() - ref_dx).abs().max()
