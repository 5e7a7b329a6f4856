"""TEST INFRASTRUCTURE -- CPU restatement (plain torch fp32) of the reference's DETR transformer layers.

Pinned by tests/golden/detr.npz, produced by oracle/gen_golden_detr.py from the UNMODIFIED reference classes
(yolov7/modeling/backbone/detr_backbone.py:140-242, which wrap torch's nn.MultiheadAttention); tests/test_detr_oracle_golden.py
re-checks this file against those vectors on every CPU run.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it.
Dropout (0.1 in the reference, active in training) is an RNG-driven regulariser: parity runs use eval mode / p = 0.
"""
import math

import torch
import torch.nn.functional as F

LN_EPS = 1e-5  # nn.LayerNorm default (detr_backbone.py:146-147)
# True: round every tensor the CUDA path stores in bf16 (activations and the packed weights) -- the yardstick for judging the 16-bit path
# against this fp32 restatement (same idea as oracle/yolox_oracle.py).  Rounding has an identity gradient, so autograd still works.
EMULATE_STORAGE = False


def _q(t):
    return t.to(torch.bfloat16).to(torch.float32) if EMULATE_STORAGE else t


def mha(query, key, value, sd, prefix, nhead, key_padding_mask=None, need_probs=False):
    """nn.MultiheadAttention.forward (detr_backbone.py:140,160-161,200-202) for seq-first inputs [L, B, E]:
    packed in_proj (q | k | v rows of in_proj_weight), q scaled by head_dim^-0.5, softmax over keys with key_padding_mask (True = ignore)
    as -inf, out_proj.  Returns [Lq, B, E]."""
    lq, b, e = query.shape
    lk = key.shape[0]
    dh = e // nhead
    w, bias = sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"]
    q = _q(F.linear(query, _q(w[:e]), bias[:e])) * (dh ** -0.5)
    k = _q(F.linear(key, _q(w[e:2 * e]), bias[e:2 * e]))
    v = _q(F.linear(value, _q(w[2 * e:]), bias[2 * e:]))
    q = q.reshape(lq, b * nhead, dh).transpose(0, 1)          # [B*H, Lq, dh]
    k = k.reshape(lk, b * nhead, dh).transpose(0, 1)
    v = v.reshape(lk, b * nhead, dh).transpose(0, 1)
    s = torch.bmm(q, k.transpose(1, 2))                        # [B*H, Lq, Lk]
    if key_padding_mask is not None:
        s = s.view(b, nhead, lq, lk).masked_fill(key_padding_mask[:, None, None, :], float("-inf")).view(b * nhead, lq, lk)
    p = torch.softmax(s, dim=-1)
    o = _q(torch.bmm(_q(p), v)).transpose(0, 1).reshape(lq, b, e)
    out = F.linear(o, _q(sd[prefix + "out_proj.weight"]), sd[prefix + "out_proj.bias"])
    return (out, p) if need_probs else out


def _ln(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], LN_EPS)


def _pos(t, pos):
    return t if pos is None else t + pos


def encoder_layer_post(src, sd, prefix, nhead, key_padding_mask=None, pos=None):
    """TransformerEncoderLayer.forward_post, detr_backbone.py:157-170 (dropout = identity)"""
    src = _q(src)
    qk = _q(_pos(src, pos))
    src = _q(_ln(_q(src + mha(qk, qk, src, sd, prefix + "self_attn.", nhead, key_padding_mask)), sd, prefix + "norm1"))
    h = _q(F.relu(F.linear(src, _q(sd[prefix + "linear1.weight"]), sd[prefix + "linear1.bias"])))
    ff = F.linear(h, _q(sd[prefix + "linear2.weight"]), sd[prefix + "linear2.bias"])
    return _q(_ln(_q(src + ff), sd, prefix + "norm2"))


def decoder_layer_post(tgt, memory, sd, prefix, nhead, memory_key_padding_mask=None, pos=None, query_pos=None):
    """TransformerDecoderLayer.forward_post, detr_backbone.py:221-242 (dropout = identity)"""
    tgt, memory = _q(tgt), _q(memory)
    qk = _q(_pos(tgt, query_pos))
    tgt = _q(_ln(_q(tgt + mha(qk, qk, tgt, sd, prefix + "self_attn.", nhead)), sd, prefix + "norm1"))
    tgt = _q(_ln(_q(tgt + mha(_q(_pos(tgt, query_pos)), _q(_pos(memory, pos)), memory, sd, prefix + "multihead_attn.", nhead, memory_key_padding_mask)), sd,
                 prefix + "norm2"))
    h = _q(F.relu(F.linear(tgt, _q(sd[prefix + "linear1.weight"]), sd[prefix + "linear1.bias"])))
    ff = F.linear(h, _q(sd[prefix + "linear2.weight"]), sd[prefix + "linear2.bias"])
    return _q(_ln(_q(tgt + ff), sd, prefix + "norm3"))


def attention_core(q, k, v, key_padding_mask=None, scale=None):
    """softmax(q k^T * scale + mask) v for [B, H, L, dh] tensors: the part of nn.MultiheadAttention between in_proj and out_proj"""
    scale = q.shape[-1] ** -0.5 if scale is None else scale
    s = torch.matmul(q, k.transpose(-1, -2)) * scale
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    return torch.matmul(torch.softmax(s, -1), v)


def layer_state_dict(kind, d_model, ffn, seed=0):
    """parameters of one layer under the reference's names (nn.MultiheadAttention: in_proj_weight [3E,E], in_proj_bias, out_proj.*)"""
    g = torch.Generator().manual_seed(seed)

    def rn(*s, std):
        return torch.randn(*s, generator=g) * std

    sd = {}
    for att in (["self_attn"] if kind == "encoder" else ["self_attn", "multihead_attn"]):
        sd[att + ".in_proj_weight"] = rn(3 * d_model, d_model, std=d_model ** -0.5)
        sd[att + ".in_proj_bias"] = rn(3 * d_model, std=0.1)
        sd[att + ".out_proj.weight"] = rn(d_model, d_model, std=d_model ** -0.5)
        sd[att + ".out_proj.bias"] = rn(d_model, std=0.1)
    sd["linear1.weight"], sd["linear1.bias"] = rn(ffn, d_model, std=d_model ** -0.5), rn(ffn, std=0.1)
    sd["linear2.weight"], sd["linear2.bias"] = rn(d_model, ffn, std=ffn ** -0.5), rn(d_model, std=0.1)
    for n in (["norm1", "norm2"] if kind == "encoder" else ["norm1", "norm2", "norm3"]):
        sd[n + ".weight"] = torch.rand(d_model, generator=g) + 0.5
        sd[n + ".bias"] = rn(d_model, std=0.1)
    return sd
