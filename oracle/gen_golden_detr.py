"""TEST INFRASTRUCTURE -- generates tests/golden/detr.npz from the UNMODIFIED reference transformer layers
(yolov7/modeling/backbone/detr_backbone.py imported through oracle/ref_shim.py; detectron2.utils.comm is stubbed).
Run in the build container:   python -m oracle.gen_golden_detr"""
import importlib
import os
import sys
import types

import numpy as np
import torch

from . import detr_oracle as dto
from . import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "detr.npz")


def load_reference():
    ref_shim.install()
    comm = types.ModuleType("detectron2.utils.comm")
    comm.is_main_process, comm.get_world_size, comm.synchronize = (lambda: True), (lambda: 1), (lambda: None)
    u = types.ModuleType("detectron2.utils")
    u.__path__ = []
    u.comm = comm
    sys.modules["detectron2.utils"], sys.modules["detectron2.utils.comm"] = u, comm
    return importlib.import_module("yolov7.modeling.backbone.detr_backbone")


def _np(t):
    return t.detach().cpu().numpy()


def main():
    mod = load_reference()
    out = {}
    d, nhead, ffn, b = 64, 2, 128, 3
    g = torch.Generator().manual_seed(1)
    # encoder layer: L = 150 tokens (one full and one partial 128-row tile), ragged key padding
    L = 150
    enc = mod.TransformerEncoderLayer(d, nhead, dim_feedforward=ffn, dropout=0.0)
    sd = dto.layer_state_dict("encoder", d, ffn, seed=2)
    enc.load_state_dict(sd, strict=True)
    enc.eval()
    src = torch.randn(L, b, d, generator=g, requires_grad=True)
    pos = torch.randn(L, b, d, generator=g)
    mask = torch.zeros(b, L, dtype=torch.bool)
    mask[1, 100:] = True
    mask[2, 17:40] = True
    y = enc(src, src_key_padding_mask=mask, pos=pos)
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)
    out.update(enc_src=_np(src), enc_pos=_np(pos), enc_mask=_np(mask), enc_out=_np(y), enc_gout=_np(gy), enc_gsrc=_np(src.grad))
    for k, p in enc.named_parameters():
        out["enc_grad/" + k] = _np(p.grad)
    # decoder layer: 20 queries against the 150-token memory
    dec = mod.TransformerDecoderLayer(d, nhead, dim_feedforward=ffn, dropout=0.0)
    sdd = dto.layer_state_dict("decoder", d, ffn, seed=3)
    dec.load_state_dict(sdd, strict=True)
    dec.eval()
    tgt = torch.randn(20, b, d, generator=g)
    qpos = torch.randn(20, b, d, generator=g)
    mem = torch.randn(L, b, d, generator=g)
    z = dec(tgt, mem, memory_key_padding_mask=mask, pos=pos, query_pos=qpos)
    out.update(dec_tgt=_np(tgt), dec_qpos=_np(qpos), dec_mem=_np(mem), dec_out=_np(z))
    # attention probabilities of the encoder's MultiheadAttention (need_weights averages over heads)
    qk = src.detach() + pos
    ao, aw = enc.self_attn(qk, qk, value=src.detach(), key_padding_mask=mask)
    out.update(att_out=_np(ao), att_weights_mean=_np(aw))
    out["dims"] = np.array([d, nhead, ffn, b, L])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "%.1f KB" % (os.path.getsize(OUT) / 1e3))


if __name__ == "__main__":
    main()
