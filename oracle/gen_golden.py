"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by EXECUTING THE REFERENCE (unmodified files from
/root/reference, imported through oracle/ref_shim.py) on seeded inputs.  Run in the build container only:

    python -m oracle.gen_golden

The fixtures are small (a few MB in total) and committed; the GPU box never sees /root/reference.  Each fixture
stores the inputs it was computed from (so nothing depends on RNG reproducibility) and the reference's outputs.
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle import yolox_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def _set_bn(mods):
    for m in mods.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eps, m.momentum = orc.BN_EPS, orc.BN_MOMENTUM


def trained_like_outputs(head_mod, labels, size, num_classes, seed, noise=0.15):
    """[B,A,5+C] decoded head outputs in the 'trained-like' regime of SURVEY.md par.8d(2b): anchors within 2.5 strides
    of a GT centre predict a jittered copy of that GT with confident logits; the rest are background."""
    g = torch.Generator().manual_seed(seed)
    hw = [(size // s, size // s) for s in orc.STRIDES]
    xs, ys, ss = orc.anchor_grid(hw)
    a = xs.numel()
    bsz = labels.shape[0]
    out = torch.zeros(bsz, a, 5 + num_classes)
    xc, yc = (xs + 0.5) * ss, (ys + 0.5) * ss
    out[..., 0] = xc
    out[..., 1] = yc
    out[..., 2:4] = (ss * 4)[None, :, None] * torch.exp(torch.randn(bsz, a, 2, generator=g) * 0.3)
    out[..., 4:] = torch.randn(bsz, a, 1 + num_classes, generator=g) - 4.0
    for b in range(bsz):
        n = int((labels[b].sum(1) > 0).sum())
        for i in range(n):
            c, cx, cy, w, h = labels[b, i].tolist()
            near = ((xc - cx).abs() < 2.5 * ss) & ((yc - cy).abs() < 2.5 * ss)
            k = int(near.sum())
            if k == 0:
                continue
            jit = 1 + torch.randn(k, 4, generator=g) * noise
            out[b, near, :4] = torch.tensor([cx, cy, w, h]) * jit
            out[b, near, 4] = torch.randn(k, generator=g) + 2.0
            out[b, near, 5 + int(c)] = torch.randn(k, generator=g) + 2.0
    return out


def gen_blocks(mods):
    """module-level outputs of the conv building blocks (train + eval), tiny shapes"""
    _, _, _, _, wr, ck = mods
    torch.manual_seed(11)
    res = {}
    x = torch.randn(2, 16, 12, 12)
    blocks = {
        "baseconv3": wr.BaseConv(16, 24, 3, 1),
        "baseconv3s2": wr.BaseConv(16, 24, 3, 2),
        "bottleneck": wr.Bottleneck(16, 16, True, 1.0),
        "csp": wr.CSPLayer(16, 32, n=2),
        "spp": wr.SPPBottleneck(16, 16),
        "focus": wr.Focus(4, 8, 3),
    }
    for name, m in blocks.items():
        _set_bn(m)
        for p in m.parameters():
            p.data = p.data.to(torch.bfloat16).float()
        for n_, b_ in m.named_buffers():
            if "running_mean" in n_:
                b_.normal_(0, 0.2)
            if "running_var" in n_:
                b_.uniform_(0.5, 1.5)
        inp = torch.randn(2, 4, 12, 12) if name == "focus" else x
        sd0 = {k: v.clone() for k, v in m.state_dict().items()}
        m.eval()
        y_eval = m(inp)
        m.train()
        y_train = m(inp)
        res[name + ".in"] = _np(inp)
        res[name + ".eval"] = _np(y_eval)
        res[name + ".train"] = _np(y_train)
        for k, v in sd0.items():
            res[f"{name}.sd0.{k}"] = _np(v)
        for k, v in m.state_dict().items():
            if "running" in k:
                res[f"{name}.sd1.{k}"] = _np(v)
    # eval-mode BN folding oracle (utils/checkpoint.py:11-43)
    bc = blocks["baseconv3"]
    fused = ck.fuse_conv_and_bn(bc.conv, bc.bn)
    res["fuse.weight"] = _np(fused.weight)
    res["fuse.bias"] = _np(fused.bias)
    np.savez_compressed(os.path.join(OUT, "blocks.npz"), **res)


def gen_box_losses(mods):
    boxes = mods[0]
    g = torch.Generator().manual_seed(21)
    n = 256
    pred = torch.cat([torch.rand(n, 2, generator=g) * 600, torch.exp(torch.rand(n, 2, generator=g) * 4 + 1)], 1)
    tgt = pred * (1 + torch.randn(n, 4, generator=g) * 0.2)
    tgt[:, 2:] = tgt[:, 2:].abs() + 1
    tgt[:16, :2] += 500  # some disjoint pairs
    res = {"pred": _np(pred), "target": _np(tgt)}
    for lt in ("iou", "giou"):
        p = pred.clone().requires_grad_(True)
        l = boxes.IOUloss(reduction="none", loss_type=lt)(p, tgt)
        l.sum().backward()
        res[f"iouloss.{lt}"] = _np(l)
        res[f"iouloss.{lt}.grad"] = _np(p.grad)
    for it in ("giou", "diou", "ciou"):
        p = pred.clone().requires_grad_(True)
        l = boxes.IOUlossV6(box_format="xywh", iou_type=it)(p.T, tgt)
        l.sum().backward()
        res[f"v6.{it}"] = _np(l)
        res[f"v6.{it}.grad"] = _np(p.grad)
    res["pairwise_iou"] = _np(boxes.bboxes_iou(tgt[:20], pred[:64], False))
    np.savez_compressed(os.path.join(OUT, "box_losses.npz"), **res)


def gen_simota(mods, size=256, num_classes=80):
    """get_assignments + get_losses of the reference head on [B,A,85] outputs in both regimes (+ empty image)"""
    hd = mods[3]
    head = hd.YOLOXHead(num_classes, width=0.5)
    head.train()
    hw = [(size // s, size // s) for s in orc.STRIDES]
    xs, ys, ss = orc.anchor_grid(hw)
    res = {"size": np.int64(size)}
    for case, (seed, max_gt, regime) in {"trained": (31, 12, "trained"), "init": (32, 8, "init"), "crowd": (33, 40, "trained")}.items():
        _, labels = orc.synthetic_batch(4, size, seed, max_gt=max_gt, empty_every=4)
        if regime == "trained":
            out = trained_like_outputs(head, labels, size, num_classes, seed + 100)
        else:
            g = torch.Generator().manual_seed(seed + 100)
            raw = [torch.randn(4, 5 + num_classes, h, w, generator=g) * 0.5 for (h, w) in hw]
            for r in raw:
                r[:, 4:] -= 4.6
            out = orc.decode_train(raw)
        out = out.detach().clone().requires_grad_(True)
        x_shifts = [x.view(1, -1) for x in torch.split(xs, [h * w for h, w in hw])]
        y_shifts = [y.view(1, -1) for y in torch.split(ys, [h * w for h, w in hw])]
        strides = [s.view(1, -1) for s in torch.split(ss, [h * w for h, w in hw])]
        # record per-image assignments by intercepting get_assignments
        rec = []
        orig = head.get_assignments

        def spy(*a, **k):
            r = orig(*a, **k)
            rec.append((a[0], [t.clone() if torch.is_tensor(t) else t for t in r]))
            return r

        head.get_assignments = spy
        loss, iou5, lobj, lcls, l1, ratio = head.get_losses(None, x_shifts, y_shifts, strides, labels, out, [], dtype=torch.float32)
        head.get_assignments = orig
        loss.backward()
        res[f"{case}.outputs"] = _np(out).astype(np.float32)
        res[f"{case}.labels"] = _np(labels)
        res[f"{case}.losses"] = np.array([float(loss), float(iou5), float(lobj), float(lcls), float(ratio)], dtype=np.float64)
        res[f"{case}.grad"] = _np(out.grad)
        for b, (cls_m, fg, ious, gti, nfg) in rec:
            res[f"{case}.b{b}.fg_mask"] = _np(fg)
            res[f"{case}.b{b}.matched_gt"] = _np(gti)
            res[f"{case}.b{b}.matched_cls"] = _np(cls_m)
            res[f"{case}.b{b}.matched_iou"] = _np(ious)
        res[f"{case}.images_with_gt"] = np.array([b for b, _ in rec], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "simota.npz"), **res)


from yolov7_d2_b200.synth import clustered_predictions  # noqa: E402,F401


def gen_nms(mods, num_classes=80):
    boxes = mods[0]
    pred = clustered_predictions(3, 2100, num_classes, 41)
    pred[2, :, 4] = 0.0  # one image with nothing above the confidence threshold -> None
    ref = boxes.postprocess(pred.clone(), num_classes, 0.001, 0.65)
    res = {"pred": _np(pred).astype(np.float32), "conf_thre": np.float64(0.001), "nms_thre": np.float64(0.65)}
    for i, d in enumerate(ref):
        res[f"det{i}"] = _np(d) if d is not None else np.zeros((0, 7), np.float32)
    # higher confidence threshold: fewer candidates (coordinate-trick path in torchvision when numel <= 4000)
    ref2 = boxes.postprocess(pred.clone(), num_classes, 0.3, 0.45)
    for i, d in enumerate(ref2):
        res[f"det_hi{i}"] = _np(d) if d is not None else np.zeros((0, 7), np.float32)
    np.savez_compressed(os.path.join(OUT, "nms.npz"), **res)


def gen_model(mods, width=0.25, depth=0.33, size=128, num_classes=80):
    """whole reference model (CSPDarknet -> YOLOPAFPN -> YOLOXHead) fwd+bwd, narrow width to keep the fixture small"""
    _, dk, pa, hd, _, _ = mods
    torch.manual_seed(51)
    bb = dk.CSPDarknet(depth, width, out_features=["dark3", "dark4", "dark5"])
    neck = pa.YOLOPAFPN(depth=depth, width=width, in_features=["dark3", "dark4", "dark5"])
    head = hd.YOLOXHead(num_classes, width=width)
    for m in (bb, neck, head):
        _set_bn(m)
        for p in m.parameters():
            p.data = p.data.to(torch.bfloat16).float()
    head.initialize_biases(1e-2)
    sd = {}
    for pre, m in (("backbone.", bb), ("neck.", neck), ("head.", head)):
        for k, v in m.state_dict().items():
            sd[pre + k] = v.clone()
    images, labels = orc.synthetic_batch(2, size, 52, max_gt=6)
    x = images.float()
    for m in (bb, neck, head):
        m.train()
    loss, iou5, lobj, lcls, l1, ratio = head(neck(bb(x)), labels, x)
    loss.backward()
    res = {"images": _np(images), "labels": _np(labels),
           "losses": np.array([float(loss), float(iou5), float(lobj), float(lcls), float(ratio)], dtype=np.float64)}
    for k, v in sd.items():
        if v.dtype == torch.float32 and v.dim() == 4:
            res["sd." + k] = _np(v.to(torch.bfloat16).view(torch.int16))  # bf16-exact weights, stored as raw bits
        else:
            res["sd." + k] = _np(v)
    grads = {}
    for pre, m in (("backbone.", bb), ("neck.", neck), ("head.", head)):
        for k, p in m.named_parameters():
            grads[pre + k] = p.grad
    for k in ("backbone.stem.conv.conv.weight", "backbone.dark3.1.m.0.conv2.conv.weight", "backbone.dark5.2.conv3.bn.weight",
              "neck.C3_p3.conv3.conv.weight", "head.cls_preds.0.bias", "head.obj_preds.2.weight", "head.reg_convs.1.0.bn.bias"):
        res["grad." + k] = _np(grads[k])
    res["grad_norms_keys"] = np.array(sorted(grads.keys()))
    res["grad_norms"] = np.array([float(grads[k].norm()) for k in sorted(grads.keys())], dtype=np.float64)
    res["bn.backbone.dark2.0.bn.running_mean"] = _np(bb.dark2[0].bn.running_mean)
    res["bn.backbone.dark2.0.bn.running_var"] = _np(bb.dark2[0].bn.running_var)
    for m in (bb, neck, head):
        m.eval()
    with torch.no_grad():
        ev = head(neck(bb(x)))
    res["eval_out"] = _np(ev)
    det = mods[0].postprocess(ev.clone(), num_classes, 0.001, 0.65)
    res["eval_num_det"] = np.array([0 if d is None else d.shape[0] for d in det])
    np.savez_compressed(os.path.join(OUT, "model_w025.npz"), **res)


def main():
    os.makedirs(OUT, exist_ok=True)
    mods = ref_shim.load()
    torch.set_num_threads(8)
    gen_blocks(mods)
    gen_box_losses(mods)
    gen_simota(mods)
    gen_nms(mods)
    gen_model(mods)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
