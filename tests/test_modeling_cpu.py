"""Host-side marshalling of yolov7_d2_b200.modeling that needs no GPU: label packing (yolox.py:150-162)."""
import types

import torch

from yolov7_d2_b200 import modeling


def test_stage_labels_matches_reference_packing():
    fake = types.SimpleNamespace(max_boxes_num=5)
    g = torch.Generator().manual_seed(0)
    inputs, ref = [], torch.zeros(4, 5, 5)
    for k, n in enumerate((3, 0, 7, 1)):  # 7 > max_boxes_num: truncated like labels[:max_boxes_num]
        xy = torch.rand(n, 2, generator=g) * 100
        wh = torch.rand(n, 2, generator=g) * 50 + 1
        boxes = torch.cat([xy, xy + wh], 1)
        cls = torch.randint(0, 80, (n,), generator=g)
        inst = types.SimpleNamespace(gt_boxes=types.SimpleNamespace(tensor=boxes), gt_classes=cls)
        inputs.append({"instances": inst} if k != 3 else {"targets": inst})
        m = min(n, 5)
        ref[k, :m, 0] = cls[:m].float()
        ref[k, :m, 1] = (boxes[:m, 0] + boxes[:m, 2]) / 2
        ref[k, :m, 2] = (boxes[:m, 1] + boxes[:m, 3]) / 2
        ref[k, :m, 3] = boxes[:m, 2] - boxes[:m, 0]
        ref[k, :m, 4] = boxes[:m, 3] - boxes[:m, 1]
    dst = torch.full((4, 5, 5), 7.0)
    modeling.YOLOX._stage_labels(fake, inputs, dst)
    assert torch.equal(dst, ref)
    inputs.append({"image": None})  # an input without annotations
    dst = torch.full((5, 5, 5), 7.0)
    modeling.YOLOX._stage_labels(fake, inputs, dst)
    assert torch.equal(dst[:4], ref) and float(dst[4].abs().sum()) == 0.0
