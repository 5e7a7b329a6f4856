"""Host logic of the DETR encoder layer's training path without a GPU: `_EncoderLayerFn` (which tensor goes into which kernel, which slice
of the packed q|k|v buffer, which residual joins where, which gradient lands in which parameter) is executed with every kernel wrapper
replaced by its torch fp32 equivalent, and must reproduce the reference layer's autograd (tests/golden/detr.npz) to 1e-5.
The stand-ins live in this test only -- the product has no CPU path."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import detr_oracle as dto

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "detr.npz")


def _sl(x):
    if isinstance(x, tuple):
        t, o, c = x
        return t[..., o:o + c]
    return x


class TorchKernels:
    """same method names / argument conventions as yolov7_d2_b200.detr._Kernels"""

    def add(self, a, b):
        return a + b

    def linear(self, x, w, bias, out=None, out_off=0, residual=None, relu=False):
        y = F.linear(_sl(x), w.detach(), bias.detach())
        if residual is not None:
            y = y + residual
        if relu:
            y = F.relu(y)
        if out is None:
            return y
        out[..., out_off:out_off + y.shape[-1]] = y
        return out

    def pack2(self, w):
        return w.detach(), w.detach()

    def layernorm_train(self, x, w, b):
        stats = torch.stack([x.mean(-1).flatten(), (x.var(-1, unbiased=False) + 1e-5).rsqrt().flatten()], 1)
        return F.layer_norm(x, (x.shape[-1],), w.detach(), b.detach(), 1e-5), stats

    def layernorm_bwd(self, dy, x, stats, w):
        with torch.enable_grad():
            xr, wr = x.clone().requires_grad_(True), w.detach().clone().requires_grad_(True)
            br = torch.zeros_like(wr).requires_grad_(True)
            F.layer_norm(xr, (x.shape[-1],), wr, br, 1e-5).backward(dy)
        return xr.grad, wr.grad, br.grad

    def dgrad(self, dz, w, cin, addend=None):
        y = _sl(dz) @ w
        return y + addend if addend is not None else y

    def dgrad_relu(self, dz, w, h):
        du = (dz @ w) * (h > 0)
        return du, du.sum((0, 1, 2))

    def wgrad(self, x, dz, out):
        out.copy_(torch.einsum("bhlo,bhli->oi", _sl(dz), _sl(x)))

    def colsum(self, dz, out):
        out.copy_(_sl(dz).sum((0, 1, 2)))

    @staticmethod
    def _heads(t, heads):
        b, _, l, e = t.shape
        return t.view(b, l, heads, e // heads).permute(0, 2, 1, 3)

    @staticmethod
    def _attn_drop(qt, kt, heads, p_drop, seed):
        if p_drop <= 0:
            return None
        return dto.attention_dropout_multiplier(seed, qt.shape[0], heads, qt.shape[2], kt.shape[2], p_drop)

    def attention_train(self, q, k, v, mask, heads, p_drop=0.0, seed=0):
        qt, kt = _sl(q), _sl(k)
        b, _, lq, e = qt.shape
        o = dto.attention_core(self._heads(qt, heads), self._heads(kt, heads), self._heads(_sl(v), heads), mask.bool() if mask is not None else None,
                               attn_drop=self._attn_drop(qt, kt, heads, p_drop, seed))
        return o.permute(0, 2, 1, 3).reshape(b, 1, lq, e), None

    def attention_bwd(self, q, k, v, out, dout, mask, heads, lse, dq, dk, dv, p_drop=0.0, seed=0):
        with torch.enable_grad():
            qt, kt, vt = (_sl(t).clone().requires_grad_(True) for t in (q, k, v))
            b, _, lq, e = qt.shape
            o = dto.attention_core(self._heads(qt, heads), self._heads(kt, heads), self._heads(vt, heads), mask.bool() if mask is not None else None,
                                   attn_drop=self._attn_drop(qt, kt, heads, p_drop, seed))
            o.permute(0, 2, 1, 3).reshape(b, 1, lq, e).backward(dout)
        for (t, off, c), g in zip((dq, dk, dv), (qt.grad, kt.grad, vt.grad)):
            t[..., off:off + c] = g

    def dropout(self, x, p_drop, seed, residual=None, scale=1.0):
        b, _, l, c = x.shape
        y = x * dto.dropout_multiplier(seed, (b, l, c), p_drop).view(b, 1, l, c) * scale
        return y + residual if residual is not None else y


@pytest.fixture()
def detr_fp32(monkeypatch):
    """import the module without the CUDA library and keep its internal buffers in fp32"""
    from yolov7_d2_b200 import detr

    monkeypatch.setattr(detr, "_bl", lambda t: t.detach().permute(1, 0, 2).float().contiguous().unsqueeze(1))
    monkeypatch.setattr(detr, "_lb", lambda t: t.squeeze(1).permute(1, 0, 2).float().contiguous())
    monkeypatch.setattr(torch, "bfloat16", torch.float32)  # the Function allocates its packed buffers with torch.bfloat16
    return detr


def test_encoder_layer_training_wiring(detr_fp32):
    detr = detr_fp32
    gold = np.load(GOLD, allow_pickle=False)
    d, nhead, ffn, b, L = (int(v) for v in gold["dims"])
    layer = type("Layer", (), {})()
    layer.k, layer.d_model, layer.nhead = TorchKernels(), d, nhead
    layer._dropout_state = lambda n: (0.0, (0,) * n)
    sd = dto.layer_state_dict("encoder", d, ffn, seed=2)
    params = [sd[n].clone().requires_grad_(True) for n in detr._EncoderLayerFn.NAMES]
    src = torch.tensor(gold["enc_src"]).requires_grad_(True)
    pos = torch.tensor(gold["enc_pos"]).requires_grad_(True)
    out = detr._EncoderLayerFn.apply(layer, src, pos, torch.tensor(gold["enc_mask"]).to(torch.uint8), *params)

    def close(a, ref, what):
        ref = torch.as_tensor(np.asarray(ref))
        err = (a - ref).abs().max().item()
        assert err <= 1e-5 * ref.abs().max().item(), f"{what}: {err:.3e}"

    close(out.detach(), gold["enc_out"], "output")
    out.backward(torch.tensor(gold["enc_gout"]))
    close(src.grad, gold["enc_gsrc"], "src gradient")
    for n, p in zip(detr._EncoderLayerFn.NAMES, params):
        close(p.grad, gold["enc_grad/" + n], n)
    # the positional embedding receives the gradient of the query / key path only
    sdr = {"l." + k: v for k, v in sd.items()}
    s2, p2 = torch.tensor(gold["enc_src"]), torch.tensor(gold["enc_pos"]).requires_grad_(True)
    dto.encoder_layer_post(s2, sdr, "l.", nhead, torch.tensor(gold["enc_mask"]), p2).backward(torch.tensor(gold["enc_gout"]))
    close(pos.grad, p2.grad, "pos gradient")


def test_decoder_layer_training_wiring(detr_fp32):
    """`_DecoderLayerFn` against the autograd of oracle.detr_oracle.decoder_layer_post (whose forward is pinned to the reference layer)"""
    detr = detr_fp32
    gold = np.load(GOLD, allow_pickle=False)
    d, nhead, ffn, b, L = (int(v) for v in gold["dims"])
    layer = type("Layer", (), {})()
    layer.k, layer.d_model, layer.nhead = TorchKernels(), d, nhead
    layer._dropout_state = lambda n: (0.0, (0,) * n)
    sd = dto.layer_state_dict("decoder", d, ffn, seed=3)
    params = [sd[n].clone().requires_grad_(True) for n in detr._DecoderLayerFn.NAMES]
    names = ("dec_tgt", "dec_mem", "enc_pos", "dec_qpos")
    ours = [torch.tensor(gold[k]).requires_grad_(True) for k in names]
    mem_mask = torch.tensor(gold["enc_mask"])
    tgt_mask = torch.zeros(b, ours[0].shape[0], dtype=torch.bool)
    tgt_mask[1, 15:] = True
    out = detr._DecoderLayerFn.apply(layer, ours[0], ours[1], ours[2], ours[3], tgt_mask.to(torch.uint8), mem_mask.to(torch.uint8), *params)
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(4))
    out.backward(gout)
    # reference: autograd of the oracle (self-attention key padding is not an argument of decoder_layer_post: restate the block here)
    refs = [torch.tensor(gold[k]).requires_grad_(True) for k in names]
    sdr = {"l." + k: v.clone().requires_grad_(True) for k, v in sd.items()}
    tgt, mem, pos, qpos = refs
    qk = tgt + qpos
    t1 = dto._ln(tgt + dto.mha(qk, qk, tgt, sdr, "l.self_attn.", nhead, tgt_mask), sdr, "l.norm1")
    t2 = dto._ln(t1 + dto.mha(t1 + qpos, mem + pos, mem, sdr, "l.multihead_attn.", nhead, mem_mask), sdr, "l.norm2")
    ff = F.linear(F.relu(F.linear(t2, sdr["l.linear1.weight"], sdr["l.linear1.bias"])), sdr["l.linear2.weight"], sdr["l.linear2.bias"])
    ref_out = dto._ln(t2 + ff, sdr, "l.norm3")
    ref_out.backward(gout)

    def close(a, ref, what):
        err = (a - ref).abs().max().item()
        assert err <= 2e-5 * ref.abs().max().item(), f"{what}: {err:.3e}"

    close(out.detach(), ref_out.detach(), "output")
    for k, a, r in zip(names, ours, refs):
        close(a.grad, r.grad, k + " gradient")
    for n, p in zip(detr._DecoderLayerFn.NAMES, params):
        close(p.grad, sdr["l." + n].grad, n)


P_DROP = 0.1


def test_encoder_layer_training_wiring_with_dropout(detr_fp32):
    """dropout = 0.1 (the reference's default, detr_backbone.py:132,140-152): the four masks of one step (attention probabilities, dropout1, FFN
    dropout, dropout2) enter the forward and are re-applied in the right places of the backward -- against the autograd of the oracle layer with the
    same explicit masks (oracle.detr_oracle: the kernels' counter-based hash restated in torch)"""
    detr = detr_fp32
    gold = np.load(GOLD, allow_pickle=False)
    d, nhead, ffn, b, L = (int(v) for v in gold["dims"])
    seeds = (11, 22, 33, 44)
    layer = type("Layer", (), {})()
    layer.k, layer.d_model, layer.nhead = TorchKernels(), d, nhead
    layer._dropout_state = lambda n: (P_DROP, seeds)
    sd = dto.layer_state_dict("encoder", d, ffn, seed=2)
    params = [sd[n].clone().requires_grad_(True) for n in detr._EncoderLayerFn.NAMES]
    src = torch.tensor(gold["enc_src"]).requires_grad_(True)
    pos = torch.tensor(gold["enc_pos"]).requires_grad_(True)
    mask = torch.tensor(gold["enc_mask"])
    out = detr._EncoderLayerFn.apply(layer, src, pos, mask.to(torch.uint8), *params)
    gout = torch.tensor(gold["enc_gout"])
    out.backward(gout)
    drop = (dto.attention_dropout_multiplier(seeds[0], b, nhead, L, L, P_DROP), dto.dropout_multiplier(seeds[1], (b, L, d), P_DROP),
            dto.dropout_multiplier(seeds[2], (b, L, ffn), P_DROP), dto.dropout_multiplier(seeds[3], (b, L, d), P_DROP))
    sdr = {"l." + k: v.clone().requires_grad_(True) for k, v in sd.items()}
    s2, p2 = torch.tensor(gold["enc_src"]).requires_grad_(True), torch.tensor(gold["enc_pos"]).requires_grad_(True)
    ref = dto.encoder_layer_post(s2, sdr, "l.", nhead, mask, p2, drop=drop)
    ref.backward(gout)
    plain = dto.encoder_layer_post(torch.tensor(gold["enc_src"]), {k: v.detach() for k, v in sdr.items()}, "l.", nhead, mask, torch.tensor(gold["enc_pos"]))
    assert (ref.detach() - plain).abs().max() > 1e-2, "the masks must change the result"

    def close(a, r, what):
        err = (a - r).abs().max().item()
        assert err <= 2e-5 * r.abs().max().item() + 1e-7, f"{what}: {err:.3e}"

    close(out.detach(), ref.detach(), "output")
    close(src.grad, s2.grad, "src gradient")
    close(pos.grad, p2.grad, "pos gradient")
    for n, p in zip(detr._EncoderLayerFn.NAMES, params):
        close(p.grad, sdr["l." + n].grad, n)


def test_decoder_layer_training_wiring_with_dropout(detr_fp32):
    detr = detr_fp32
    gold = np.load(GOLD, allow_pickle=False)
    d, nhead, ffn, b, L = (int(v) for v in gold["dims"])
    seeds = (5, 6, 7, 8, 9, 10)
    layer = type("Layer", (), {})()
    layer.k, layer.d_model, layer.nhead = TorchKernels(), d, nhead
    layer._dropout_state = lambda n: (P_DROP, seeds)
    sd = dto.layer_state_dict("decoder", d, ffn, seed=3)
    params = [sd[n].clone().requires_grad_(True) for n in detr._DecoderLayerFn.NAMES]
    names = ("dec_tgt", "dec_mem", "enc_pos", "dec_qpos")
    ours = [torch.tensor(gold[k]).requires_grad_(True) for k in names]
    mem_mask = torch.tensor(gold["enc_mask"])
    out = detr._DecoderLayerFn.apply(layer, ours[0], ours[1], ours[2], ours[3], None, mem_mask.to(torch.uint8), *params)
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(4))
    out.backward(gout)
    lq, lk = ours[0].shape[0], ours[1].shape[0]
    drop = (dto.attention_dropout_multiplier(seeds[0], b, nhead, lq, lq, P_DROP), dto.dropout_multiplier(seeds[1], (b, lq, d), P_DROP),
            dto.attention_dropout_multiplier(seeds[2], b, nhead, lq, lk, P_DROP), dto.dropout_multiplier(seeds[3], (b, lq, d), P_DROP),
            dto.dropout_multiplier(seeds[4], (b, lq, ffn), P_DROP), dto.dropout_multiplier(seeds[5], (b, lq, d), P_DROP))
    refs = [torch.tensor(gold[k]).requires_grad_(True) for k in names]
    sdr = {"l." + k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = dto.decoder_layer_post(refs[0], refs[1], sdr, "l.", nhead, mem_mask, refs[2], refs[3], drop=drop)
    ref.backward(gout)

    def close(a, r, what):
        err = (a - r).abs().max().item()
        assert err <= 2e-5 * r.abs().max().item() + 1e-7, f"{what}: {err:.3e}"

    close(out.detach(), ref.detach(), "output")
    for k, a, r in zip(names, ours, refs):
        close(a.grad, r.grad, k + " gradient")
    for n, p in zip(detr._DecoderLayerFn.NAMES, params):
        close(p.grad, sdr["l." + n].grad, n)


def test_restated_dropout_hash_properties():
    """oracle.detr_oracle.{dropout_multiplier, attention_dropout_multiplier}: deterministic in the seed, keep rate 1 - p, kept values 1 / (1 - p),
    different seeds / heads / rows give different masks (the kernels' own masks are compared with these bit for bit on the GPU)"""
    m1, m2, m3 = dto.dropout_multiplier(9, (4, 64, 256), 0.1), dto.dropout_multiplier(9, (4, 64, 256), 0.1), dto.dropout_multiplier(10, (4, 64, 256), 0.1)
    assert torch.equal(m1, m2) and not torch.equal(m1, m3)
    assert abs(float((m1 > 0).float().mean()) - 0.9) < 0.01
    assert set(m1.unique().tolist()) == {0.0, float(torch.tensor(1.0) / (torch.tensor(1.0) - torch.tensor(0.1)))}
    a = dto.attention_dropout_multiplier(3, 2, 8, 100, 120, 0.1)
    assert abs(float((a > 0).float().mean()) - 0.9) < 0.01
    assert not torch.equal(a[0, 0], a[0, 1]) and not torch.equal(a[0, 0], a[1, 0]) and not torch.equal(a[0, 0, 0], a[0, 0, 1])
    # p = 0.5: threshold exactly half of the 24-bit range
    assert abs(float((dto.dropout_multiplier(1, (1, 1000, 64), 0.5) > 0).float().mean()) - 0.5) < 0.01
