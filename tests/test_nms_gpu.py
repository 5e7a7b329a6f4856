"""GPU parity of the batched NMS post-processing: bit-exact detections (boxes, scores, classes, order) against the
reference-generated golden fixture and against the CPU oracle at full size."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import yolox_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_nms(capi, pred, conf, thr, dev, mutate=0):
    L = capi.lib()
    b, a, ch = pred.shape
    p = pred.to(dev).contiguous()
    ws = torch.empty(L.yb200_nms_workspace(b, a), dtype=torch.uint8, device=dev)
    det = torch.full((b, a, 7), float("nan"), device=dev)
    cnt = torch.empty(b, dtype=torch.int32, device=dev)
    capi.check(L.yb200_postprocess_nms(capi.ptr(p), b, a, ch - 5, ctypes.c_float(conf), ctypes.c_float(thr), mutate, capi.ptr(ws), capi.ptr(det),
                                       capi.ptr(cnt), capi.stream_ptr()), "postprocess_nms")
    torch.cuda.synchronize()
    return [det[i, :int(cnt[i])].cpu() for i in range(b)], p.cpu()


@pytest.mark.parametrize("tag,conf,thr", [("det", 0.001, 0.65), ("det_hi", 0.3, 0.45)])
def test_nms_vs_reference_golden(cuda, tag, conf, thr):
    from yolov7_d2_b200 import capi

    g = np.load(os.path.join(GOLD, "nms.npz"))
    pred = torch.from_numpy(g["pred"])
    dets, _ = run_nms(capi, pred, conf, thr, cuda)
    for i, d in enumerate(dets):
        ref = torch.from_numpy(g[f"{tag}{i}"])
        assert d.shape == ref.shape, (i, d.shape, ref.shape)
        assert torch.equal(d, ref), f"image {i}"


def test_nms_full_size_vs_oracle(cuda):
    from oracle.gen_golden import clustered_predictions
    from yolov7_d2_b200 import capi

    pred = clustered_predictions(4, 8400, 80, 91)
    pred[3, :, 4] = 0.0005  # nothing passes the confidence filter
    dets, mutated = run_nms(capi, pred, 0.001, 0.65, cuda, mutate=1)
    ref = orc.postprocess(pred, 80, 0.001, 0.65)
    for i, (d, r) in enumerate(zip(dets, ref)):
        if r is None:
            assert d.shape[0] == 0
            continue
        assert d.shape == r.shape and torch.equal(d, r), f"image {i}: {d.shape} vs {r.shape}"
    # the reference rewrites prediction[..., :4] to corners in place (boxes.py:177)
    xyxy = torch.stack([pred[..., 0] - pred[..., 2] / 2, pred[..., 1] - pred[..., 3] / 2, pred[..., 0] + pred[..., 2] / 2,
                        pred[..., 1] + pred[..., 3] / 2], -1)
    assert torch.equal(mutated[..., :4], xyxy) and torch.equal(mutated[..., 4:], pred[..., 4:])


def test_nms_single_class_chain(cuda):
    """adversarial: one class, heavily overlapping chain -> long sequential suppression"""
    from yolov7_d2_b200 import capi

    n = 3000
    pred = torch.zeros(1, n, 85)
    pred[0, :, 0] = torch.arange(n) * 0.7 + 50
    pred[0, :, 1] = 100
    pred[0, :, 2:4] = 40
    pred[0, :, 4] = torch.linspace(0.9, 0.1, n)
    pred[0, :, 5] = 0.8
    dets, _ = run_nms(capi, pred, 0.001, 0.5, cuda)
    ref = orc.postprocess(pred, 80, 0.001, 0.5)
    assert torch.equal(dets[0], ref[0])


def test_nms_tied_scores_vs_reference_golden(cuda):
    """tests/golden/nms_ties.npz (reference `postprocess` on colliding scores, oracle/gen_golden_nms_ties.py): the device order for equal
    scores is the stable one (= reference on CUDA / <= 1000 candidates: `small`, bit-exact); postprocess(tie_order="torch_cpu_sort")
    reproduces the reference's CPU `_batched_nms_vanilla` order (`big`, bit-exact)."""
    from yolov7_d2_b200 import capi
    from yolov7_d2_b200.modeling import postprocess

    g = np.load(os.path.join(GOLD, "nms_ties.npz"))
    dets, _ = run_nms(capi, torch.from_numpy(g["small.pred"]), 0.001, 0.65, cuda)
    for i, d in enumerate(dets):
        assert torch.equal(d, torch.from_numpy(g[f"small.det{i}"])), f"small image {i}"
    big = torch.from_numpy(g["big.pred"]).to(cuda)
    got = postprocess(big.clone(), 80, 0.001, 0.65, tie_order="torch_cpu_sort")
    canon = postprocess(big.clone(), 80, 0.001, 0.65)
    ref_canon = orc.postprocess(torch.from_numpy(g["big.pred"]), 80, 0.001, 0.65)
    for i, d in enumerate(got):
        ref = torch.from_numpy(g[f"big.det{i}"])
        assert torch.equal(d.cpu(), ref), f"big image {i}: tie_order=torch_cpu_sort differs from the reference"
        assert torch.equal(canon[i].cpu(), ref_canon[i]), f"big image {i}: canonical order differs from the oracle"
        assert not torch.equal(canon[i].cpu(), ref)
