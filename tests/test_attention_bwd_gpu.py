"""GPU parity of the attention backward (yb200_attention_bwd) against torch autograd of the oracle's attention_core on the same
bf16-rounded q, k, v, dout.  P and dS are rounded to bf16 before the tensor-core products => 3e-2 of each gradient's max."""
import ctypes
import os

import pytest
import torch

from oracle import detr_oracle as dto

pytestmark = [pytest.mark.gpu]

CASES = [(2, 2, 150, 150, True), (1, 1, 128, 128, False), (2, 8, 100, 300, True), (1, 4, 260, 70, False)]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "B%d_H%d_q%d_k%d_%s" % (c[0], c[1], c[2], c[3], "mask" if c[4] else "nomask"))
def test_attention_backward(cuda, case):
    from yolov7_d2_b200 import capi

    L = capi.lib()
    b, heads, lq, lk, masked = case
    e = heads * 32
    g = torch.Generator().manual_seed(lq * 3 + lk)
    mk = lambda l, s=1.0: (torch.randn(b, 1, l, e, generator=g) * s).to(cuda).to(torch.bfloat16)
    q, k, v, dout = mk(lq, 1.5), mk(lk, 1.5), mk(lk), mk(lq)
    mask = None
    if masked:
        mask = torch.zeros(b, lk, dtype=torch.uint8)
        mask[0, lk - lk // 3:] = 1
        if b > 1:
            mask[1, 1:max(2, lk // 5)] = 1
        mask = mask.to(cuda)
    scale = 32 ** -0.5
    out = torch.empty_like(q)
    lse = torch.empty(b, heads, lq, device=cuda)
    A = capi.act
    qa, ka, va, oa, da = A(q), A(k), A(v), A(out), A(dout)
    capi.check(L.yb200_attention_fwd(ctypes.byref(qa), ctypes.byref(ka), ctypes.byref(va), capi.ptr(mask), ctypes.c_float(scale), ctypes.byref(oa), capi.ptr(lse),
                                     capi.stream_ptr()), "fwd")
    dq, dk, dv = (torch.full_like(t, float("nan")) for t in (q, k, v))
    dqa, dka, dva = A(dq), A(dk), A(dv)
    ws = torch.empty(int(L.yb200_attention_bwd_workspace(ctypes.byref(qa))), dtype=torch.uint8, device=cuda)
    capi.check(L.yb200_attention_bwd(ctypes.byref(qa), ctypes.byref(ka), ctypes.byref(va), ctypes.byref(oa), ctypes.byref(da), capi.ptr(mask), ctypes.c_float(scale),
                                     capi.ptr(lse), ctypes.byref(dqa), ctypes.byref(dka), ctypes.byref(dva), capi.ptr(ws), capi.stream_ptr()), "bwd")
    torch.cuda.synchronize()

    def hf(t, l):
        return t.float().view(b, l, heads, 32).permute(0, 2, 1, 3)

    qr, kr, vr = (hf(t, l).clone().requires_grad_(True) for t, l in ((q, lq), (k, lk), (v, lk)))
    ref = dto.attention_core(qr, kr, vr, mask.bool() if masked else None, scale)
    ref.backward(hf(dout, lq))
    for name, got, r, l in (("dq", dq, qr.grad, lq), ("dk", dk, kr.grad, lk), ("dv", dv, vr.grad, lk)):
        gh = hf(got, l)
        assert torch.isfinite(gh).all(), name
        err = (gh - r).abs().max().item()
        assert err <= 3e-2 * r.abs().max().item(), f"{name}: max err {err:.4e} vs max |ref| {r.abs().max().item():.3f}"
    # bit-reproducible (no atomics)
    dq2, dk2, dv2 = (torch.empty_like(t) for t in (q, k, v))
    a2 = [A(t) for t in (dq2, dk2, dv2)]
    capi.check(L.yb200_attention_bwd(ctypes.byref(qa), ctypes.byref(ka), ctypes.byref(va), ctypes.byref(oa), ctypes.byref(da), capi.ptr(mask), ctypes.c_float(scale),
                                     capi.ptr(lse), ctypes.byref(a2[0]), ctypes.byref(a2[1]), ctypes.byref(a2[2]), capi.ptr(ws), capi.stream_ptr()), "bwd")
    assert torch.equal(dq, dq2) and torch.equal(dk, dk2) and torch.equal(dv, dv2)
