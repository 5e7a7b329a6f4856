"""oracle/sparseinst_oracle.py against vectors produced by the unmodified reference BaseIAMDecoder (tests/golden/sparseinst.npz,
oracle/gen_golden_sparseinst.py).  fp32 CPU on both sides: 2e-5 of the tensor's max."""
import os

import numpy as np
import torch

from oracle import sparseinst_oracle as sio

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sparseinst.npz")


def test_decoder_forward_matches_reference():
    gold = np.load(GOLD, allow_pickle=False)
    dim, nm, kd, nc, convs, cin = (int(v) for v in gold["dims"])
    sd = sio.decoder_state_dict(5, in_channels=cin, dim=dim, num_masks=nm, kernel_dim=kd, num_classes=nc, num_convs=convs)
    out = sio.decoder_forward(torch.tensor(gold["feat"]), sd, num_convs=convs)
    for k in ("pred_logits", "pred_masks", "pred_scores", "pred_kernel", "iam"):
        ref = torch.tensor(gold[k])
        err = (out[k] - ref).abs().max().item()
        assert err <= 2e-5 * ref.abs().max().item(), f"{k}: {err:.3e}"
    assert out["pred_masks"].shape == (2, nm, 24, 40)


def test_group_decoder_forward_matches_reference():
    """GroupIAMDecoder (decoder_sparseinst.py:172-250) against tests/golden/sparseinst_group.npz"""
    gold = np.load(GOLD.replace("sparseinst.npz", "sparseinst_group.npz"), allow_pickle=False)
    dim, nm, kd, nc, convs, cin, groups = (int(v) for v in gold["dims"])
    sd = sio.decoder_state_dict(7, in_channels=cin, dim=dim, num_masks=nm, kernel_dim=kd, num_classes=nc, num_convs=convs, groups=groups)
    out = sio.decoder_forward(torch.tensor(gold["feat"]), sd, num_convs=convs, groups=groups)
    for k in ("pred_logits", "pred_masks", "pred_scores", "pred_kernel", "iam"):
        ref = torch.tensor(gold[k])
        err = (out[k] - ref).abs().max().item()
        assert err <= 2e-5 * ref.abs().max().item(), f"{k}: {err:.3e}"
    assert out["iam"].shape[1] == nm * groups and out["pred_logits"].shape == (2, nm, nc)
