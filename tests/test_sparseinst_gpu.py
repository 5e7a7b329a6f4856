"""SparseInst IAM decoder (yolov7_d2_b200.sparseinst.BaseIAMDecoder) against the outputs of the unmodified reference decoder
(tests/golden/sparseinst.npz).  bf16 storage of every intermediate vs the fp32 reference: 5e-2 of each tensor's max, correlation > 0.999."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import sparseinst_oracle as sio

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sparseinst.npz")


def ns(**kw):
    return types.SimpleNamespace(**kw)


def _check(got, ref, what, tol=5e-2):
    got, ref = got.float().cpu(), torch.as_tensor(np.asarray(ref)).float()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert torch.isfinite(got).all(), what
    err = (got - ref).abs().max().item()
    cos = torch.dot(got.flatten(), ref.flatten()) / (got.norm() * ref.norm())
    assert err <= tol * ref.abs().max().item() and cos > 0.999, f"{what}: max err {err:.4f} (max |ref| {ref.abs().max().item():.3f}), cos {cos:.5f}"


def _cfg(dim, nm, kd, nc, convs, cin, iam=False):
    return ns(MODEL=ns(SPARSE_INST=ns(ENCODER=ns(NUM_CHANNELS=cin), DECODER=ns(SCALE_FACTOR=2.0, OUTPUT_IAM=iam, NUM_MASKS=nm, KERNEL_DIM=kd, NUM_CLASSES=nc,
                                                                                   INST=ns(DIM=dim, CONVS=convs), MASK=ns(DIM=dim, CONVS=convs)))))


def test_decoder_matches_reference(cuda):
    from yolov7_d2_b200.sparseinst import BaseIAMDecoder

    gold = np.load(GOLD, allow_pickle=False)
    dim, nm, kd, nc, convs, cin = (int(v) for v in gold["dims"])
    dec = BaseIAMDecoder(_cfg(dim, nm, kd, nc, convs, cin, iam=True))
    sd = sio.decoder_state_dict(5, in_channels=cin, dim=dim, num_masks=nm, kernel_dim=kd, num_classes=nc, num_convs=convs)
    dec.load_state_dict({k: v.to(cuda) for k, v in sd.items()}, strict=True)
    out = dec(torch.tensor(gold["feat"]).to(cuda))
    _check(dec.last["iam"].permute(0, 3, 1, 2), gold["iam"], "instance activation maps")
    _check(dec.last["pred_kernel"], gold["pred_kernel"], "mask kernels")
    _check(out["pred_logits"], gold["pred_logits"], "class logits")
    _check(out["pred_scores"], gold["pred_scores"], "objectness")
    _check(out["pred_masks"], gold["pred_masks"], "masks")
    assert out["pred_iam"].shape == (2, nm, 24, 40)


def test_decoder_full_size_against_oracle(cuda):
    """the shipped configuration (256+2 input channels -> padded to 272, 100 masks -> 112, kernel dim 128, 80 classes) on a 40x40 map"""
    from yolov7_d2_b200.sparseinst import BaseIAMDecoder

    dec = BaseIAMDecoder(_cfg(256, 100, 128, 80, 4, 256))
    sd = sio.decoder_state_dict(9)
    dec.load_state_dict({k: v.to(cuda) for k, v in sd.items()}, strict=True)
    feat = torch.randn(2, 256, 40, 40, generator=torch.Generator().manual_seed(10))
    torch.set_num_threads(min(16, torch.get_num_threads()))
    ref = sio.decoder_forward(feat, sd)
    out = dec(feat.to(cuda))
    _check(out["pred_logits"], ref["pred_logits"], "class logits")
    _check(out["pred_scores"], ref["pred_scores"], "objectness")
    _check(dec.last["masks_lowres"], ref["masks_lowres"], "masks before up-sampling")
    _check(out["pred_masks"], ref["pred_masks"], "masks")


def test_group_decoder_matches_reference(cuda):
    """GroupIAMDecoder (grouped IAM convolution, N*G maps, fc + ReLU: decoder_sparseinst.py:172-250) against the unmodified reference
    (tests/golden/sparseinst_group.npz)"""
    from yolov7_d2_b200.sparseinst import GroupIAMDecoder

    gold = np.load(GOLD.replace("sparseinst.npz", "sparseinst_group.npz"), allow_pickle=False)
    dim, nm, kd, nc, convs, cin, groups = (int(v) for v in gold["dims"])
    cfg = _cfg(dim, nm, kd, nc, convs, cin, iam=True)
    cfg.MODEL.SPARSE_INST.DECODER.GROUPS = groups
    dec = GroupIAMDecoder(cfg)
    sd = sio.decoder_state_dict(7, in_channels=cin, dim=dim, num_masks=nm, kernel_dim=kd, num_classes=nc, num_convs=convs, groups=groups)
    dec.load_state_dict({k: v.to(cuda) for k, v in sd.items()}, strict=True)
    out = dec(torch.tensor(gold["feat"]).to(cuda))
    _check(dec.last["iam"].permute(0, 3, 1, 2), gold["iam"], "grouped instance activation maps")
    _check(dec.last["pred_kernel"], gold["pred_kernel"], "mask kernels")
    _check(out["pred_logits"], gold["pred_logits"], "class logits")
    _check(out["pred_scores"], gold["pred_scores"], "objectness")
    _check(out["pred_masks"], gold["pred_masks"], "masks")
    assert out["pred_iam"].shape == (2, nm * groups, 24, 40)


def test_group_decoder_full_size_against_oracle(cuda):
    """the shipped Group-IAM configuration: 4 groups x 100 masks, dim 256 -> 1024-wide instance features, 40x40 map"""
    from yolov7_d2_b200.sparseinst import GroupIAMDecoder

    cfg = _cfg(256, 100, 128, 80, 4, 256)
    cfg.MODEL.SPARSE_INST.DECODER.GROUPS = 4
    dec = GroupIAMDecoder(cfg)
    sd = sio.decoder_state_dict(11, groups=4)
    dec.load_state_dict({k: v.to(cuda) for k, v in sd.items()}, strict=True)
    feat = torch.randn(2, 256, 40, 40, generator=torch.Generator().manual_seed(12))
    torch.set_num_threads(min(16, torch.get_num_threads()))
    ref = sio.decoder_forward(feat, sd, groups=4)
    out = dec(feat.to(cuda))
    _check(out["pred_logits"], ref["pred_logits"], "class logits")
    _check(out["pred_scores"], ref["pred_scores"], "objectness")
    _check(out["pred_masks"], ref["pred_masks"], "masks")


@pytest.mark.parametrize("shape", [(3, 5, 7), (200, 80, 80), (1, 1, 1)], ids=str)
def test_bilinear_x2_matches_torch(cuda, shape):
    """yb200_upsample_bilinear2x_f32 == F.interpolate(scale_factor=2, mode="bilinear", align_corners=False) (decoder_sparseinst.py:148-153)"""
    import ctypes

    import torch.nn.functional as F

    from yolov7_d2_b200 import capi

    planes, h, w = shape
    x = torch.randn(planes, h, w, generator=torch.Generator().manual_seed(9)).to(cuda)
    out = torch.empty(planes, 2 * h, 2 * w, device=cuda)
    capi.check(capi.lib().yb200_upsample_bilinear2x_f32(capi.ptr(x), capi.ptr(out), ctypes.c_int64(planes), h, w, capi.stream_ptr()), "bilinear")
    ref = F.interpolate(x[None], scale_factor=2, mode="bilinear", align_corners=False)[0]
    assert torch.allclose(out, ref, rtol=1e-6, atol=1e-6), float((out - ref).abs().max())
