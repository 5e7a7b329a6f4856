"""TEST INFRASTRUCTURE -- tests/golden/nms_ties.npz: the UNMODIFIED reference `postprocess` (yolov7/utils/boxes.py:171-210, imported
through oracle/ref_shim.py) on prediction sets with deliberately TIED scores (obj and class confidences drawn from coarse grids, so
obj * cls collides across classes and inside a class).  Run in the build container only:  python -m oracle.gen_golden_nms_ties

Why a separate fixture: with tied scores the reference's output order is decided by torch's sort implementation, not by its own code --
  * > 1000 candidates on the CPU: torchvision `_batched_nms_vanilla` ends with `scores[keep].sort(descending=True)` (UNSTABLE: libstdc++
    introsort on (key, index) pairs; ties come out in an input-dependent permutation);
  * <= 1000 candidates on the CPU, and every YOLOX-sized call on CUDA (numel <= 100 000): `_batched_nms_coordinate_trick`, i.e. the order
    `nms` returns = a STABLE descending sort (ties: lower index first).
The fixture records which strategy produced each image.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def tied_predictions(batch, anchors, num_classes, seed, levels=4):
    """clustered boxes (as synth.clustered_predictions) whose obj / class confidences take `levels` distinct values each"""
    from yolov7_d2_b200.synth import clustered_predictions
    pred = clustered_predictions(batch, anchors, num_classes, seed)
    g = torch.Generator().manual_seed(seed + 1000)
    obj_grid = torch.tensor([0.25, 0.5, 0.75, 1.0])[:levels]
    cls_grid = torch.tensor([0.9, 0.6, 0.45, 0.3])[:levels]  # 0.5 * 0.9 == 0.75 * 0.6 == 1.0 * 0.45: ties across different (obj, cls) pairs
    arg = pred[..., 5:].argmax(-1)
    pred[..., 4] = obj_grid[torch.randint(0, levels, (batch, anchors), generator=g)]
    pred[..., 5:] = 0.01
    conf = cls_grid[torch.randint(0, levels, (batch, anchors), generator=g)]
    pred.scatter_(2, (arg + 5).unsqueeze(-1), conf.unsqueeze(-1))
    return pred


def main():
    boxes_mod = ref_shim.load()[0]
    res = {}
    for tag, anchors, seed in (("big", 2100, 5), ("small", 600, 6)):
        pred = tied_predictions(2, anchors, 80, seed)
        ref = boxes_mod.postprocess(pred.clone(), 80, 0.001, 0.65)
        res[f"{tag}.pred"] = pred.numpy().astype(np.float32)
        for i, d in enumerate(ref):
            res[f"{tag}.det{i}"] = d.numpy()
            sc = d[:, 4] * d[:, 5]
            ties = int((sc[1:] == sc[:-1]).sum())
            print(tag, i, "kept", d.shape[0], "adjacent equal scores", ties, "strategy", "vanilla" if anchors * 4 > 4000 else "coordinate_trick")
            assert ties > 10
        res[f"{tag}.strategy"] = np.array("vanilla" if anchors * 4 > 4000 else "coordinate_trick")
    np.savez_compressed(os.path.join(OUT, "nms_ties.npz"), **res)


if __name__ == "__main__":
    main()
