"""Dropout of the DETR transformer layers (detr_backbone.py:132-152, 200-214: nn.MultiheadAttention(dropout=0.1) on the attention probabilities,
nn.Dropout on the residual branches and inside the FFN).  The kernels draw their masks from a counter-based hash of (seed, element index)
(csrc/attention.cu) that oracle/detr_oracle.py restates in torch, so every check is "the same computation given the same mask":
  * yb200_dropout == the oracle's multiplier, bit for bit (on ones), and residual / scale handling;
  * yb200_attention_fwd_dropout / _bwd_dropout == autograd of the oracle's attention_core with the oracle's [B, H, Lq, Lk] multiplier;
  * a training step of the encoder / decoder layer == autograd of the oracle layer with the explicit masks (16-bit storage tolerance);
  * the keep rate is 1 - p (statistical), p = 0 and eval mode take the dropout-free path."""
import ctypes

import pytest
import torch

from oracle import detr_oracle as dto

pytestmark = [pytest.mark.gpu]
P = 0.1


def test_dropout_kernel_matches_the_restated_hash(cuda):
    from yolov7_d2_b200 import capi

    L = capi.lib()
    b, l, c, seed = 3, 50, 256, 12345
    ones = torch.ones(b, 1, l, c, dtype=torch.bfloat16, device=cuda)
    out = torch.empty_like(ones)
    xa, oa = capi.act(ones), capi.act(out)
    capi.check(L.yb200_dropout(ctypes.byref(xa), None, ctypes.byref(oa), ctypes.c_float(P), ctypes.c_uint32(seed), ctypes.c_float(1.0), capi.stream_ptr()), "dropout")
    ref = dto.dropout_multiplier(seed, (b, l, c), P).view(b, 1, l, c)
    assert torch.equal(out.float().cpu() > 0, ref > 0), "keep pattern differs from the oracle's hash"
    assert torch.allclose(out.float().cpu(), ref.to(torch.bfloat16).float())
    keep = float((out > 0).float().mean())
    assert abs(keep - (1 - P)) < 0.01, keep
    # residual + x * mask * scale, on a channel slice of a wider buffer (the mask follows the logical element index, not the memory offset)
    g = torch.Generator().manual_seed(1)
    wide = torch.randn(b, 1, l, 2 * c, generator=g).to(cuda).to(torch.bfloat16)
    res = torch.randn(b, 1, l, c, generator=g).to(cuda).to(torch.bfloat16)
    out2 = torch.empty_like(res)
    xs, ra, o2 = capi.act(wide, c, c), capi.act(res), capi.act(out2)
    capi.check(L.yb200_dropout(ctypes.byref(xs), ctypes.byref(ra), ctypes.byref(o2), ctypes.c_float(P), ctypes.c_uint32(seed), ctypes.c_float(0.5), capi.stream_ptr()), "dropout")
    want = res.float().cpu() + wide[..., c:].float().cpu() * ref * 0.5
    assert torch.allclose(out2.float().cpu(), want, rtol=2 ** -7, atol=1e-2)
    # p = 0: identity
    capi.check(L.yb200_dropout(ctypes.byref(xa), None, ctypes.byref(oa), ctypes.c_float(0.0), ctypes.c_uint32(seed), ctypes.c_float(1.0), capi.stream_ptr()), "dropout")
    assert torch.equal(out, ones)


@pytest.mark.parametrize("case", [(2, 2, 150, 150, True), (2, 8, 100, 300, True), (1, 4, 260, 70, False)], ids=str)
def test_attention_dropout_forward_and_backward(cuda, case):
    from yolov7_d2_b200 import capi

    L = capi.lib()
    b, heads, lq, lk, masked = case
    e, seed = heads * 32, 777 + lq
    g = torch.Generator().manual_seed(lq * 3 + lk)
    mk = lambda l, s=1.0: (torch.randn(b, 1, l, e, generator=g) * s).to(cuda).to(torch.bfloat16)
    q, k, v, dout = mk(lq, 1.5), mk(lk, 1.5), mk(lk), mk(lq)
    mask = None
    if masked:
        mask = torch.zeros(b, lk, dtype=torch.uint8)
        mask[0, lk - lk // 3:] = 1
        mask = mask.to(cuda)
    scale = 32 ** -0.5
    out, lse = torch.empty_like(q), torch.empty(b, heads, lq, device=cuda)
    A = capi.act
    qa, ka, va, oa, da = A(q), A(k), A(v), A(out), A(dout)
    capi.check(L.yb200_attention_fwd_dropout(ctypes.byref(qa), ctypes.byref(ka), ctypes.byref(va), capi.ptr(mask), ctypes.c_float(scale), ctypes.byref(oa), capi.ptr(lse),
                                             ctypes.c_float(P), ctypes.c_uint32(seed), capi.stream_ptr()), "fwd")
    dq, dk, dv = (torch.full_like(t, float("nan")) for t in (q, k, v))
    dqa, dka, dva = A(dq), A(dk), A(dv)
    ws = torch.empty(int(L.yb200_attention_bwd_workspace(ctypes.byref(qa))), dtype=torch.uint8, device=cuda)
    capi.check(L.yb200_attention_bwd_dropout(ctypes.byref(qa), ctypes.byref(ka), ctypes.byref(va), ctypes.byref(oa), ctypes.byref(da), capi.ptr(mask), ctypes.c_float(scale),
                                             capi.ptr(lse), ctypes.byref(dqa), ctypes.byref(dka), ctypes.byref(dva), capi.ptr(ws), ctypes.c_float(P), ctypes.c_uint32(seed),
                                             capi.stream_ptr()), "bwd")
    torch.cuda.synchronize()
    hf = lambda t, l: t.float().cpu().view(b, l, heads, 32).permute(0, 2, 1, 3)
    qr, kr, vr = (hf(t, l).clone().requires_grad_(True) for t, l in ((q, lq), (k, lk), (v, lk)))
    mult = dto.attention_dropout_multiplier(seed, b, heads, lq, lk, P)
    ref = dto.attention_core(qr, kr, vr, mask.cpu().bool() if masked else None, scale, attn_drop=mult)
    plain = dto.attention_core(qr.detach(), kr.detach(), vr.detach(), mask.cpu().bool() if masked else None, scale)
    err = (hf(out, lq) - ref.detach()).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item(), f"forward with dropout: {err:.4e}"
    assert (ref.detach() - plain).abs().max() > 10 * err, "the mask must matter at this tolerance"
    ref.backward(hf(dout, lq))
    for name, got, r, l in (("dq", dq, qr.grad, lq), ("dk", dk, kr.grad, lk), ("dv", dv, vr.grad, lk)):
        gh = hf(got, l)
        assert torch.isfinite(gh).all(), name
        e_ = (gh - r).abs().max().item()
        assert e_ <= 3e-2 * r.abs().max().item(), f"{name}: max err {e_:.4e} vs max |ref| {r.abs().max().item():.3f}"
    # the log-sum-exp (and so the softmax normaliser) is the one of the UNDROPPED probabilities
    lse_plain = torch.empty_like(lse)
    capi.check(L.yb200_attention_fwd(ctypes.byref(qa), ctypes.byref(ka), ctypes.byref(va), capi.ptr(mask), ctypes.c_float(scale), ctypes.byref(oa), capi.ptr(lse_plain),
                                     capi.stream_ptr()), "fwd plain")
    assert torch.equal(lse, lse_plain)


def _layer_case(cuda, kind):
    from yolov7_d2_b200 import detr

    d, nhead, ffn, b, lsrc, lq = 256, 8, 512, 2, 90, 40
    cls = detr.TransformerEncoderLayer if kind == "encoder" else detr.TransformerDecoderLayer
    layer = cls(d, nhead, dim_feedforward=ffn, dropout=P).train()
    sd = dto.layer_state_dict(kind, d, ffn, seed=5)
    layer.load_state_dict({k: v.to(cuda) for k, v in sd.items()})
    g = torch.Generator().manual_seed(8)
    rn = lambda *s: torch.randn(*s, generator=g)
    return layer, sd, d, nhead, ffn, b, lsrc, lq, rn


def _yard(a, ref, emu, what):
    """16-bit storage yardstick as in tests/test_detr_gpu.py: error <= 2.5 x the error of the storage-emulating oracle + 2 % of the maximum"""
    e_k, e_e, m = (a - ref).abs().max().item(), (emu - ref).abs().max().item(), ref.abs().max().item()
    assert e_k <= 2.5 * e_e + 0.02 * m, f"{what}: kernel err {e_k:.4e}, 16-bit oracle err {e_e:.4e}, max {m:.3f}"


def test_encoder_layer_trains_with_dropout(cuda):
    layer, sd, d, nhead, ffn, b, L, _, rn = _layer_case(cuda, "encoder")
    seeds = (101, 202, 303, 404)
    layer._dropout_state = lambda n: (P, seeds)
    src, pos, gout = rn(L, b, d), rn(L, b, d), rn(L, b, d)
    mask = torch.zeros(b, L, dtype=torch.bool)
    mask[1, 70:] = True
    s = src.to(cuda).requires_grad_(True)
    out = layer(s, src_key_padding_mask=mask.to(cuda), pos=pos.to(cuda))
    out.backward(gout.to(cuda))
    drop = (dto.attention_dropout_multiplier(seeds[0], b, nhead, L, L, P), dto.dropout_multiplier(seeds[1], (b, L, d), P),
            dto.dropout_multiplier(seeds[2], (b, L, ffn), P), dto.dropout_multiplier(seeds[3], (b, L, d), P))
    res = {}
    for emulate in (False, True):
        dto.EMULATE_STORAGE = emulate
        try:
            sdr = {"l." + k: v.clone().requires_grad_(True) for k, v in sd.items()}
            s2 = src.clone().requires_grad_(True)
            o = dto.encoder_layer_post(s2, sdr, "l.", nhead, mask, pos, drop=drop)
            o.backward(gout)
            res[emulate] = (o.detach(), s2.grad, {k: v.grad for k, v in sdr.items()})
        finally:
            dto.EMULATE_STORAGE = False
    _yard(out.detach().float().cpu(), res[False][0], res[True][0], "output")
    _yard(s.grad.float().cpu(), res[False][1], res[True][1], "src gradient")
    for n, p in layer.named_parameters():
        _yard(p.grad.float().cpu(), res[False][2]["l." + n], res[True][2]["l." + n], n)
    # eval mode and p = 0 take the dropout-free path: two calls agree exactly
    del layer._dropout_state
    layer.eval()
    with torch.no_grad():
        a, c = layer(src.to(cuda), src_key_padding_mask=mask.to(cuda), pos=pos.to(cuda)), layer(src.to(cuda), src_key_padding_mask=mask.to(cuda), pos=pos.to(cuda))
    assert torch.equal(a, c)
    # training mode draws fresh seeds per call (torch's CPU generator): two steps differ, the same torch seed reproduces
    layer.train()
    outs = []
    for seed in (1, 2, 1):
        torch.manual_seed(seed)
        outs.append(layer(src.to(cuda).requires_grad_(True), src_key_padding_mask=mask.to(cuda), pos=pos.to(cuda)).detach())
    assert not torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_decoder_layer_trains_with_dropout(cuda):
    layer, sd, d, nhead, ffn, b, lk, lq, rn = _layer_case(cuda, "decoder")
    seeds = (11, 12, 13, 14, 15, 16)
    layer._dropout_state = lambda n: (P, seeds)
    tgt, mem, pos, qpos, gout = rn(lq, b, d), rn(lk, b, d), rn(lk, b, d), rn(lq, b, d), rn(lq, b, d)
    mask = torch.zeros(b, lk, dtype=torch.bool)
    mask[0, 60:] = True
    t, m = tgt.to(cuda).requires_grad_(True), mem.to(cuda).requires_grad_(True)
    out = layer(t, m, memory_key_padding_mask=mask.to(cuda), pos=pos.to(cuda), query_pos=qpos.to(cuda))
    out.backward(gout.to(cuda))
    drop = (dto.attention_dropout_multiplier(seeds[0], b, nhead, lq, lq, P), dto.dropout_multiplier(seeds[1], (b, lq, d), P),
            dto.attention_dropout_multiplier(seeds[2], b, nhead, lq, lk, P), dto.dropout_multiplier(seeds[3], (b, lq, d), P),
            dto.dropout_multiplier(seeds[4], (b, lq, ffn), P), dto.dropout_multiplier(seeds[5], (b, lq, d), P))
    res = {}
    for emulate in (False, True):
        dto.EMULATE_STORAGE = emulate
        try:
            sdr = {"l." + k: v.clone().requires_grad_(True) for k, v in sd.items()}
            t2, m2 = tgt.clone().requires_grad_(True), mem.clone().requires_grad_(True)
            o = dto.decoder_layer_post(t2, m2, sdr, "l.", nhead, mask, pos, qpos, drop=drop)
            o.backward(gout)
            res[emulate] = (o.detach(), t2.grad, m2.grad, {k: v.grad for k, v in sdr.items()})
        finally:
            dto.EMULATE_STORAGE = False
    _yard(out.detach().float().cpu(), res[False][0], res[True][0], "output")
    _yard(t.grad.float().cpu(), res[False][1], res[True][1], "tgt gradient")
    _yard(m.grad.float().cpu(), res[False][2], res[True][2], "memory gradient")
    for n, p in layer.named_parameters():
        _yard(p.grad.float().cpu(), res[False][3]["l." + n], res[True][3]["l." + n], n)
