"""DETR transformer layers on the B200 kernels (forward path).

Reference: yolov7/modeling/backbone/detr_backbone.py -- `TransformerEncoderLayer` :128-187 (forward_post :157-170), `TransformerDecoderLayer`
:190-279 (forward_post :221-242); both wrap torch's nn.MultiheadAttention (:140, :200-202).  The classes below keep the reference's
constructor arguments, parameter names / shapes (`self_attn.in_proj_weight` [3E,E], `linear1.weight`, `norm1.weight` ...) and the seq-first
`[L, B, E]` fp32 interface, and run

    x (+pos) -> in_proj GEMMs (bias epilogue, q|k|v packed in one [B, L, 3E] buffer) -> yb200_attention_fwd (tcgen05, streaming softmax)
      -> out_proj GEMM (+bias +residual epilogue) -> LayerNorm -> linear1 GEMM (+bias +ReLU epilogue) -> linear2 GEMM (+bias +residual) -> LayerNorm

Internally tokens are batch-first bf16 `[B, 1, L, E]` views (yb200_act).  Scope of round 1: the forward pass (inference, and the forward half
of training); the attention backward kernel is not built yet, so these modules run under no_grad and refuse inputs that require grad.
Dropout (p = 0.1 in the reference) is identity here: parity runs use eval mode / p = 0.  `attn_mask` (never passed by the
reference's DETR) and `normalize_before=True` are not supported.  There is no CPU implementation.
"""
import ctypes

import torch
import torch.nn as nn

from . import capi

LN_EPS = 1e-5


def _bl(t):
    """[L, B, E] (any float dtype, CUDA) -> contiguous bf16 [B, 1, L, E]"""
    return t.detach().permute(1, 0, 2).to(torch.bfloat16).contiguous().unsqueeze(1)


def _lb(t):
    """[B, 1, L, E] bf16 -> [L, B, E] fp32"""
    return t.squeeze(1).permute(1, 0, 2).float().contiguous()


class _Kernels:
    """thin wrappers: torch tensors in, C-ABI calls on the current stream"""

    def __init__(self):
        self.L = capi.lib()

    @staticmethod
    def _a(t, off=0, c=None):
        return capi.act(t, off, c)

    def pack(self, w):  # [out, in] fp32 -> bf16 GEMM operand
        out_f, in_f = w.shape
        wf = torch.empty(out_f, 1, in_f, dtype=torch.bfloat16, device=w.device)
        capi.check(self.L.yb200_pack_conv_weight(capi.ptr(w.detach().contiguous()), out_f, in_f, 1, out_f, in_f, capi.ptr(wf), None, capi.stream_ptr()), "pack")
        return wf

    def add(self, a, b):
        out = torch.empty_like(a)
        aa, ba, oa = self._a(a), self._a(b), self._a(out)
        capi.check(self.L.yb200_add(ctypes.byref(aa), ctypes.byref(ba), ctypes.byref(oa), capi.stream_ptr()), "add")
        return out

    def linear(self, x, w, bias, out=None, out_off=0, residual=None, relu=False):
        """out[..., out_off:out_off+N] = x W^T + bias (+ residual) (ReLU); x may be a (tensor, off, c) slice"""
        xt, xo, xc = x if isinstance(x, tuple) else (x, 0, None)
        n_out = w.shape[0]
        b, _, l, _ = xt.shape
        if out is None:
            out = torch.empty(b, 1, l, n_out, dtype=torch.bfloat16, device=xt.device)
        xa, oa = self._a(xt, xo, xc), self._a(out, out_off, n_out)
        wf = self.pack(w)
        bias = bias.detach().contiguous()
        if relu:
            capi.check(self.L.yb200_linear_relu_fwd(ctypes.byref(xa), capi.ptr(wf), capi.ptr(bias), ctypes.byref(oa), capi.stream_ptr()), "linear+relu")
        else:
            ra = self._a(residual) if residual is not None else None
            capi.check(self.L.yb200_conv2d_affine_fwd(ctypes.byref(xa), capi.ptr(wf), None, capi.ptr(bias), ctypes.byref(ra) if ra is not None else None,
                                                      ctypes.byref(oa), 1, 1, capi.stream_ptr()), "linear")
        return out

    def layernorm(self, x, weight, bias):
        y = torch.empty_like(x)
        xa, ya = self._a(x), self._a(y)
        capi.check(self.L.yb200_layernorm_fwd(ctypes.byref(xa), capi.ptr(weight.detach().contiguous()), capi.ptr(bias.detach().contiguous()), ctypes.c_float(LN_EPS),
                                              ctypes.byref(ya), None, capi.stream_ptr()), "layernorm")
        return y

    def attention(self, q, k, v, mask, heads):
        """q, k, v: (tensor, channel offset, E) slices of [B,1,L,*] buffers; mask: uint8 [B, Lk] or None"""
        qt, qo, e = q
        b, _, lq, _ = qt.shape
        out = torch.empty(b, 1, lq, e, dtype=torch.bfloat16, device=qt.device)
        qa, ka, va, oa = self._a(*q), self._a(*k), self._a(*v), self._a(out)
        capi.check(self.L.yb200_attention_fwd(ctypes.byref(qa), ctypes.byref(ka), ctypes.byref(va), capi.ptr(mask), ctypes.c_float((e // heads) ** -0.5),
                                              ctypes.byref(oa), None, capi.stream_ptr()), "attention")
        return out


def _check_inputs(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise capi.Yb200Error("DETR layers: inputs must be CUDA tensors (no CPU path)")
        if t.requires_grad and torch.is_grad_enabled():
            raise capi.Yb200Error("DETR layers: the attention backward kernel is not built yet -- run under torch.no_grad()")


def _mask_u8(mask):
    return None if mask is None else mask.to(torch.uint8).contiguous()


class _MhaParams(nn.Module):
    """parameter container with nn.MultiheadAttention's names (detr_backbone.py:140)"""

    def __init__(self, d_model, device):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d_model, d_model, device=device))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d_model, device=device))
        self.out_proj = nn.Module()
        self.out_proj.weight = nn.Parameter(torch.empty(d_model, d_model, device=device))
        self.out_proj.bias = nn.Parameter(torch.zeros(d_model, device=device))
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.kaiming_uniform_(self.out_proj.weight, a=5 ** 0.5)


class _Norm(nn.Module):
    def __init__(self, d, device):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d, device=device))
        self.bias = nn.Parameter(torch.zeros(d, device=device))


class _Lin(nn.Module):
    def __init__(self, i, o, device):
        super().__init__()
        lin = nn.Linear(i, o)
        self.weight = nn.Parameter(lin.weight.detach().to(device))
        self.bias = nn.Parameter(lin.bias.detach().to(device))


class _LayerBase(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward, dropout, activation, normalize_before, device):
        super().__init__()
        if d_model % nhead or d_model // nhead != 32:
            raise capi.Yb200Error(f"attention kernel is built for head dimension 32 (got d_model={d_model}, nhead={nhead})")
        if activation != "relu":
            raise capi.Yb200Error("only the reference's default activation (relu) is implemented")
        if normalize_before:
            raise capi.Yb200Error("normalize_before=True (forward_pre) is not implemented")
        self.d_model, self.nhead = d_model, nhead
        self.linear1 = _Lin(d_model, dim_feedforward, device)
        self.linear2 = _Lin(dim_feedforward, d_model, device)
        self.k = _Kernels()

    def _self_attention(self, x, qk, att, mask):
        """x: value source, qk: query/key source ([B,1,L,E] bf16); returns x + out_proj(attention)"""
        e, kn = self.d_model, self.k
        b, _, l, _ = x.shape
        qkv = torch.empty(b, 1, l, 3 * e, dtype=torch.bfloat16, device=x.device)
        w, bias = att.in_proj_weight, att.in_proj_bias
        kn.linear(qk, w[:2 * e], bias[:2 * e], out=qkv, out_off=0)
        kn.linear(x, w[2 * e:], bias[2 * e:], out=qkv, out_off=2 * e)
        a = kn.attention((qkv, 0, e), (qkv, e, e), (qkv, 2 * e, e), mask, self.nhead)
        return kn.linear(a, att.out_proj.weight, att.out_proj.bias, residual=x)

    def _ffn(self, x, norm):
        kn = self.k
        h = kn.linear(x, self.linear1.weight, self.linear1.bias, relu=True)
        y = kn.linear(h, self.linear2.weight, self.linear2.bias, residual=x)
        return kn.layernorm(y, norm.weight, norm.bias)


class TransformerEncoderLayer(_LayerBase):
    """detr_backbone.py:128-187"""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False, device="cuda"):
        super().__init__(d_model, nhead, dim_feedforward, dropout, activation, normalize_before, device)
        self.self_attn = _MhaParams(d_model, device)
        self.norm1, self.norm2 = _Norm(d_model, device), _Norm(d_model, device)

    @torch.no_grad()
    def forward(self, src, src_mask=None, src_key_padding_mask=None, pos=None):
        if src_mask is not None:
            raise capi.Yb200Error("attn_mask is not supported (the reference's DETR never passes one)")
        _check_inputs(src, pos)
        x = _bl(src)
        qk = x if pos is None else self.k.add(x, _bl(pos))
        y = self._self_attention(x, qk, self.self_attn, _mask_u8(src_key_padding_mask))
        x1 = self.k.layernorm(y, self.norm1.weight, self.norm1.bias)
        return _lb(self._ffn(x1, self.norm2))


class TransformerDecoderLayer(_LayerBase):
    """detr_backbone.py:190-279"""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False, device="cuda"):
        super().__init__(d_model, nhead, dim_feedforward, dropout, activation, normalize_before, device)
        self.self_attn = _MhaParams(d_model, device)
        self.multihead_attn = _MhaParams(d_model, device)
        self.norm1, self.norm2, self.norm3 = _Norm(d_model, device), _Norm(d_model, device), _Norm(d_model, device)

    @torch.no_grad()
    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None, memory_key_padding_mask=None, pos=None, query_pos=None):
        if tgt_mask is not None or memory_mask is not None:
            raise capi.Yb200Error("attn_mask is not supported (the reference's DETR never passes one)")
        _check_inputs(tgt, memory, pos, query_pos)
        kn, e = self.k, self.d_model
        x = _bl(tgt)
        qp = None if query_pos is None else _bl(query_pos)
        qk = x if qp is None else kn.add(x, qp)
        x = kn.layernorm(self._self_attention(x, qk, self.self_attn, _mask_u8(tgt_key_padding_mask)), self.norm1.weight, self.norm1.bias)
        # cross attention: queries from the decoder stream, keys / values from the encoder memory
        mem = _bl(memory)
        memk = mem if pos is None else kn.add(mem, _bl(pos))
        att = self.multihead_attn
        w, bias = att.in_proj_weight, att.in_proj_bias
        q = kn.linear(x if qp is None else kn.add(x, qp), w[:e], bias[:e])
        b, _, lk, _ = mem.shape
        kv = torch.empty(b, 1, lk, 2 * e, dtype=torch.bfloat16, device=mem.device)
        kn.linear(memk, w[e:2 * e], bias[e:2 * e], out=kv, out_off=0)
        kn.linear(mem, w[2 * e:], bias[2 * e:], out=kv, out_off=e)
        a = kn.attention((q, 0, e), (kv, 0, e), (kv, e, e), _mask_u8(memory_key_padding_mask), self.nhead)
        x = kn.layernorm(kn.linear(a, att.out_proj.weight, att.out_proj.bias, residual=x), self.norm2.weight, self.norm2.bias)
        return _lb(self._ffn(x, self.norm3))
