"""oracle/detr_oracle.py against vectors produced by the unmodified reference layers (tests/golden/detr.npz, oracle/gen_golden_detr.py).
fp32 CPU on both sides: 2e-5 of the tensor's max."""
import os

import numpy as np
import pytest
import torch

from oracle import detr_oracle as dto

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "detr.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD, allow_pickle=False)


def close(a, b, rtol=2e-5, what=""):
    a, b = torch.as_tensor(a), torch.as_tensor(np.asarray(b))
    err = (a - b).abs().max().item()
    assert err <= rtol * max(b.abs().max().item(), 1e-12), f"{what}: max err {err:.3e}"


def test_encoder_layer_forward_backward(gold):
    d, nhead, ffn, b, L = (int(v) for v in gold["dims"])
    sd = {"l." + k: v.requires_grad_(True) for k, v in dto.layer_state_dict("encoder", d, ffn, seed=2).items()}
    src = torch.tensor(gold["enc_src"]).requires_grad_(True)
    y = dto.encoder_layer_post(src, sd, "l.", nhead, torch.tensor(gold["enc_mask"]), torch.tensor(gold["enc_pos"]))
    close(y.detach(), gold["enc_out"], what="encoder output")
    y.backward(torch.tensor(gold["enc_gout"]))
    close(src.grad, gold["enc_gsrc"], what="src gradient")
    for k in gold.files:
        if k.startswith("enc_grad/"):
            close(sd["l." + k[len("enc_grad/"):]].grad, gold[k], what=k)


def test_decoder_layer_forward(gold):
    d, nhead, ffn, b, L = (int(v) for v in gold["dims"])
    sd = {"l." + k: v for k, v in dto.layer_state_dict("decoder", d, ffn, seed=3).items()}
    z = dto.decoder_layer_post(torch.tensor(gold["dec_tgt"]), torch.tensor(gold["dec_mem"]), sd, "l.", nhead, torch.tensor(gold["enc_mask"]),
                               torch.tensor(gold["enc_pos"]), torch.tensor(gold["dec_qpos"]))
    close(z, gold["dec_out"], what="decoder output")


def test_attention_probabilities_and_core(gold):
    d, nhead, ffn, b, L = (int(v) for v in gold["dims"])
    sd = {"a." + k[len("self_attn."):]: v for k, v in dto.layer_state_dict("encoder", d, ffn, seed=2).items() if k.startswith("self_attn.")}
    src, pos, mask = torch.tensor(gold["enc_src"]), torch.tensor(gold["enc_pos"]), torch.tensor(gold["enc_mask"])
    out, p = dto.mha(src + pos, src + pos, src, sd, "a.", nhead, mask, need_probs=True)
    close(out, gold["att_out"], what="attention output")
    close(p.view(b, nhead, L, L).mean(1), gold["att_weights_mean"], what="head-averaged attention weights")
    assert (p.view(b, nhead, L, L)[1, :, :, 100:] == 0).all()
    # attention_core is the same arithmetic on [B,H,L,dh] tensors
    e, dh = d, d // nhead
    w, bias = sd["a.in_proj_weight"], sd["a.in_proj_bias"]
    qk = src + pos
    q = torch.nn.functional.linear(qk, w[:e], bias[:e]).view(L, b, nhead, dh).permute(1, 2, 0, 3)
    k = torch.nn.functional.linear(qk, w[e:2 * e], bias[e:2 * e]).view(L, b, nhead, dh).permute(1, 2, 0, 3)
    v = torch.nn.functional.linear(src, w[2 * e:], bias[2 * e:]).view(L, b, nhead, dh).permute(1, 2, 0, 3)
    o = dto.attention_core(q, k, v, mask).permute(2, 0, 1, 3).reshape(L, b, e)
    close(torch.nn.functional.linear(o, sd["a.out_proj.weight"], sd["a.out_proj.bias"]), gold["att_out"], what="attention core")
