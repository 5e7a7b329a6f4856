"""Golden fixture for the L1 branch of YOLOXHead.get_losses (`use_l1 = True`, yolox_head.py:186-195, 389-429, 443-448), generated from the UNMODIFIED
reference head (oracle/ref_shim.py).  Written to tests/golden/simota_l1.npz:
    <case>.raw      [B, A, 85]  raw head outputs (the leaf: decoded boxes AND origin_preds derive from it, as reg_output / reg_output.clone())
    <case>.labels   [B, G, 5]
    <case>.losses   (total, 5*iou, obj, cls, l1, num_fg / num_gt)
    <case>.grad     d total / d raw
Run inside the build container (needs /root/reference):  python -m oracle.gen_golden_l1"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim, yolox_oracle as orc  # noqa: E402
from oracle.gen_golden import trained_like_outputs  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main(size=256, num_classes=80):
    mods = ref_shim.load()
    torch.set_num_threads(8)
    head = mods[3].YOLOXHead(num_classes, width=0.5)
    head.train()
    head.use_l1 = True
    hw = [(size // s, size // s) for s in orc.STRIDES]
    counts = [h * w for h, w in hw]
    xs, ys, ss = orc.anchor_grid(hw)
    res = {"size": np.int64(size)}
    for case, (seed, max_gt, regime) in {"trained": (41, 12, "trained"), "init": (42, 8, "init")}.items():
        _, labels = orc.synthetic_batch(4, size, seed, max_gt=max_gt, empty_every=4)
        g = torch.Generator().manual_seed(seed + 100)
        if regime == "trained":
            dec = trained_like_outputs(head, labels, size, num_classes, seed + 100).detach()
            raw = dec.clone()  # undo the decode: raw_xy = xy / s - grid, raw_wh = log(wh / s)
            raw[..., 0] = dec[..., 0] / ss - xs
            raw[..., 1] = dec[..., 1] / ss - ys
            raw[..., 2:4] = torch.log(dec[..., 2:4] / ss[:, None])
        else:
            raw = torch.randn(4, sum(counts), 5 + num_classes, generator=g) * 0.5
            raw[..., 4:] -= 4.6
        raw = raw.detach().clone().requires_grad_(True)
        grid = torch.stack((xs, ys), 1)[None]
        out = torch.cat([(raw[..., :2] + grid) * ss[None, :, None], torch.exp(raw[..., 2:4]) * ss[None, :, None], raw[..., 4:]], -1)  # yolox_head.py:226-245
        origin = [o.clone() for o in torch.split(raw[..., :4], counts, 1)]                                                         # :195 reg_output.clone()
        x_shifts = [x.view(1, -1) for x in torch.split(xs, counts)]
        y_shifts = [y.view(1, -1) for y in torch.split(ys, counts)]
        strides = [s.view(1, -1) for s in torch.split(ss, counts)]
        loss, iou5, lobj, lcls, l1, ratio = head.get_losses(None, x_shifts, y_shifts, strides, labels, out, origin, dtype=torch.float32)
        loss.backward()
        res[f"{case}.raw"] = raw.detach().numpy().astype(np.float32)
        res[f"{case}.labels"] = labels.numpy()
        res[f"{case}.losses"] = np.array([float(loss), float(iou5), float(lobj), float(lcls), float(l1), float(ratio)], dtype=np.float64)
        res[f"{case}.grad"] = raw.grad.numpy()
        print(case, res[f"{case}.losses"])
    np.savez_compressed(os.path.join(OUT, "simota_l1.npz"), **res)
    print("simota_l1.npz", os.path.getsize(os.path.join(OUT, "simota_l1.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
