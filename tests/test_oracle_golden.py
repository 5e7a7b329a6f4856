"""Pins the CPU oracle (oracle/yolox_oracle.py) to the reference: every fixture under tests/golden/ was produced by
executing the reference's own files (oracle/gen_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import yolox_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def T(a):
    return torch.from_numpy(np.asarray(a))


def block_sd(g, name, which="sd0"):
    pre = f"{name}.{which}."
    return {k[len(pre):]: T(g[k]).clone() for k in g.files if k.startswith(pre)}


@pytest.mark.parametrize("name", ["baseconv3", "baseconv3s2", "bottleneck", "csp", "spp", "focus"])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_blocks(name, mode):
    g = load("blocks.npz")
    sd = block_sd(g, name)
    sd = {"m." + k: v for k, v in sd.items()}
    x = T(g[name + ".in"])
    tr = mode == "train"
    if name == "baseconv3":
        y = orc.base_conv(x, sd, "m", 1, tr)
    elif name == "baseconv3s2":
        y = orc.base_conv(x, sd, "m", 2, tr)
    elif name == "bottleneck":
        y = orc.bottleneck(x, sd, "m", True, tr)
    elif name == "csp":
        y = orc.csp_layer(x, sd, "m", True, tr)
    elif name == "spp":
        y = orc.spp_bottleneck(x, sd, "m", tr)
    else:
        y = orc.base_conv(orc.focus(x), sd, "m.conv", 1, tr)
    ref = T(g[f"{name}.{mode}"])
    assert torch.allclose(y, ref, rtol=1e-5, atol=1e-5), (y - ref).abs().max()
    if tr:  # running statistics after one training step (momentum 0.03, unbiased variance)
        for k in g.files:
            if k.startswith(name + ".sd1."):
                got = sd["m." + k[len(name) + 5:]]
                assert torch.allclose(got, T(g[k]), rtol=1e-5, atol=1e-6), k


def test_fuse_conv_bn():
    g = load("blocks.npz")
    sd = block_sd(g, "baseconv3")
    sd.update(block_sd(g, "baseconv3", "sd1"))  # the reference folded after its training-mode forward
    w, b = orc.fuse_conv_bn(sd["conv.weight"], sd["bn.weight"], sd["bn.bias"], sd["bn.running_mean"], sd["bn.running_var"])
    assert torch.allclose(w, T(g["fuse.weight"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(b, T(g["fuse.bias"]), rtol=1e-5, atol=1e-6)


def test_box_losses():
    g = load("box_losses.npz")
    pred, tgt = T(g["pred"]), T(g["target"])
    for lt in ("iou", "giou"):
        p = pred.clone().requires_grad_(True)
        l = orc.iou_loss(p, tgt, lt)
        l.sum().backward()
        assert torch.allclose(l, T(g[f"iouloss.{lt}"]), rtol=1e-6, atol=1e-6)
        assert torch.allclose(p.grad, T(g[f"iouloss.{lt}.grad"]), rtol=1e-5, atol=1e-7)
    for it in ("giou", "diou", "ciou"):
        p = pred.clone().requires_grad_(True)
        l = orc.iou_loss_v6(p, tgt, it)
        l.sum().backward()
        assert torch.allclose(l, T(g[f"v6.{it}"]), rtol=1e-6, atol=1e-6)
        assert torch.allclose(p.grad, T(g[f"v6.{it}.grad"]), rtol=1e-5, atol=1e-7)
    assert torch.equal(orc.bboxes_iou_cxcywh(tgt[:20], pred[:64]), T(g["pairwise_iou"]))


@pytest.mark.parametrize("case", ["trained", "init", "crowd"])
def test_simota_and_losses(case):
    """bit-exact assignment indices / classes and matching IoUs; losses and d loss / d outputs to fp32 round-off"""
    g = load("simota.npz")
    size = int(g["size"])
    out = T(g[f"{case}.outputs"]).clone().requires_grad_(True)
    labels = T(g[f"{case}.labels"])
    xs, ys, ss = orc.anchor_grid([(size // s, size // s) for s in orc.STRIDES])
    total, iou5, lobj, lcls, ratio, assigns = orc.yolox_losses(out, labels, xs, ys, ss, return_assign=True)
    total.backward()
    got = np.array([float(total), float(iou5), float(lobj), float(lcls), float(ratio)])
    assert np.allclose(got, g[f"{case}.losses"], rtol=1e-6, atol=1e-6), (got, g[f"{case}.losses"])
    assert torch.allclose(out.grad, T(g[f"{case}.grad"]), rtol=1e-5, atol=1e-8)
    with_gt = set(int(b) for b in g[f"{case}.images_with_gt"])
    assert with_gt and len(with_gt) < labels.shape[0], "fixture must contain images with and without boxes"
    for b, (fg, mgt, mcls, miou) in enumerate(assigns):
        if b not in with_gt:
            assert not fg.any()
            continue
        assert torch.equal(fg, T(g[f"{case}.b{b}.fg_mask"])), f"fg_mask image {b}"
        assert torch.equal(mgt, T(g[f"{case}.b{b}.matched_gt"])), f"matched_gt image {b}"
        assert torch.equal(mcls, T(g[f"{case}.b{b}.matched_cls"])), f"matched_cls image {b}"
        assert torch.equal(miou, T(g[f"{case}.b{b}.matched_iou"])), f"matched_iou image {b}"


@pytest.mark.parametrize("case", ["trained", "init"])
def test_l1_branch(case):
    """`use_l1` (yolox_head.py:389-429, 443-448): fixture from the reference head with origin_preds (oracle/gen_golden_l1.py); the leaf is the RAW head
    output, from which both the decoded boxes and origin_preds derive"""
    g = load("simota_l1.npz")
    size = int(g["size"])
    raw = T(g[f"{case}.raw"]).clone().requires_grad_(True)
    labels = T(g[f"{case}.labels"])
    xs, ys, ss = orc.anchor_grid([(size // s, size // s) for s in orc.STRIDES])
    grid = torch.stack((xs, ys), 1)[None]
    out = torch.cat([(raw[..., :2] + grid) * ss[None, :, None], torch.exp(raw[..., 2:4]) * ss[None, :, None], raw[..., 4:]], -1)
    total, iou5, lobj, lcls, l1, ratio = orc.yolox_losses(out, labels, xs, ys, ss, origin_preds=raw[..., :4])
    total.backward()
    got = np.array([float(total), float(iou5), float(lobj), float(lcls), float(l1), float(ratio)])
    assert np.allclose(got, g[f"{case}.losses"], rtol=1e-6, atol=1e-6), (got, g[f"{case}.losses"])
    assert float(l1) > 0.1
    assert torch.allclose(raw.grad, T(g[f"{case}.grad"]), rtol=1e-5, atol=1e-8)


def test_simota_exercises_hard_branches():
    """the fixtures cover dynamic k > 1 and anchors contested by several ground truths"""
    g = load("simota.npz")
    size = int(g["size"])
    xs, ys, ss = orc.anchor_grid([(size // s, size // s) for s in orc.STRIDES])
    out, labels = T(g["crowd.outputs"]), T(g["crowd.labels"])
    n = int((labels[0].sum(1) > 0).sum())
    fg, mgt, _, _ = orc.simota_assign(labels[0, :n, 1:5], labels[0, :n, 0], out[0, :, :4], out[0, :, 5:], out[0, :, 4], xs, ys, ss)
    counts = torch.bincount(mgt, minlength=n)
    assert counts.max() > 1, "no gt with dynamic k > 1"


@pytest.mark.parametrize("tag,conf,thr", [("det", 0.001, 0.65), ("det_hi", 0.3, 0.45)])
def test_postprocess_nms(tag, conf, thr):
    g = load("nms.npz")
    pred = T(g["pred"])
    dets = orc.postprocess(pred, 80, conf, thr)
    for i, d in enumerate(dets):
        ref = T(g[f"{tag}{i}"])
        if d is None:
            assert ref.shape[0] == 0
            continue
        assert d.shape == ref.shape, (i, d.shape, ref.shape)
        assert torch.equal(d, ref), f"image {i}: detections differ"
    assert any(d is None for d in dets) and any(d is not None and d.shape[0] > 100 for d in dets)


def _tie_groups_equal(d, ref):
    """same rows, same score sequence; rows may only be permuted inside runs of bit-identical scores"""
    if d.shape != ref.shape:
        return False
    sd, sr = d[:, 4] * d[:, 5], ref[:, 4] * ref[:, 5]
    if not torch.equal(sd, sr):
        return False
    bounds = [0] + (torch.nonzero(sr[1:] != sr[:-1]).flatten() + 1).tolist() + [len(sr)]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        if sorted(map(tuple, d[lo:hi].tolist())) != sorted(map(tuple, ref[lo:hi].tolist())):
            return False
    return True


def test_postprocess_nms_tied_scores():
    """tests/golden/nms_ties.npz: reference `postprocess` on inputs whose scores collide (oracle/gen_golden_nms_ties.py).
    `small` (<= 1000 candidates: torchvision's coordinate-trick strategy = the order of `nms`, a STABLE sort -- also what the reference
    does on CUDA): the oracle's canonical order must match bit-exactly.  `big` (> 1000 candidates on the CPU: `_batched_nms_vanilla`,
    whose last line is an UNSTABLE torch.sort): tie_order="torch_cpu_sort" must match bit-exactly; the canonical order is the same
    detections with permutations only inside runs of equal scores."""
    g = load("nms_ties.npz")
    assert str(g["small.strategy"]) == "coordinate_trick" and str(g["big.strategy"]) == "vanilla"
    for tag, order in (("small", "stable"), ("big", "torch_cpu_sort")):
        dets = orc.postprocess(T(g[f"{tag}.pred"]), 80, 0.001, 0.65, tie_order=order)
        for i, d in enumerate(dets):
            ref = T(g[f"{tag}.det{i}"])
            sc = ref[:, 4] * ref[:, 5]
            assert int((sc[1:] == sc[:-1]).sum()) > 10, "fixture has no ties"
            assert d.shape == ref.shape and torch.equal(d, ref), f"{tag} image {i}: order differs from the reference under tie_order={order}"
    canon = orc.postprocess(T(g["big.pred"]), 80, 0.001, 0.65)
    for i, d in enumerate(canon):
        ref = T(g[f"big.det{i}"])
        assert not torch.equal(d, ref), "fixture does not exercise the unstable sort"
        assert _tie_groups_equal(d, ref), f"big image {i}: canonical order differs by more than a permutation of equal scores"


def golden_model_sd(g):
    sd = {}
    for k in g.files:
        if k.startswith("sd."):
            v = g[k]
            sd[k[3:]] = T(v).view(torch.bfloat16).float() if v.dtype == np.int16 else T(v).clone()
    return sd


def test_whole_model_fwd_bwd():
    """reference CSPDarknet+YOLOPAFPN+YOLOXHead (width 0.25) training step: losses, parameter gradients, BN running stats"""
    g = load("model_w025.npz")
    sd = golden_model_sd(g)
    params = [k for k, v in sd.items() if v.dtype == torch.float32 and "running" not in k]
    for k in params:
        sd[k].requires_grad_(True)
    images, labels = T(g["images"]).float(), T(g["labels"])
    total, iou5, lobj, lcls, ratio, _ = orc.yolox_forward_train(images, labels, sd)
    total.backward()
    got = np.array([float(total), float(iou5), float(lobj), float(lcls), float(ratio)])
    assert np.allclose(got, g["losses"], rtol=1e-5, atol=1e-5), (got, g["losses"])
    keys = [str(k) for k in g["grad_norms_keys"]]
    norms = np.array([float(sd[k].grad.norm()) for k in keys])
    assert np.allclose(norms, g["grad_norms"], rtol=2e-4, atol=1e-6), np.abs(norms / np.maximum(g["grad_norms"], 1e-12) - 1).max()
    for k in g.files:
        if k.startswith("grad."):
            ref = T(g[k])
            assert torch.allclose(sd[k[5:]].grad, ref, rtol=1e-3, atol=2e-4 * float(ref.abs().max()) + 1e-9), k  # fp32 summation-order noise through ~60 layers
    for k in ("backbone.dark2.0.bn.running_mean", "backbone.dark2.0.bn.running_var"):
        assert torch.allclose(sd[k], T(g["bn." + k]), rtol=1e-5, atol=1e-6), k


def test_whole_model_eval():
    g = load("model_w025.npz")
    sd = golden_model_sd(g)
    # the golden eval pass ran after one training step: replay it so the running statistics match
    images, labels = T(g["images"]).float(), T(g["labels"])
    with torch.no_grad():
        orc.yolox_forward_train(images, labels, sd)
        out = orc.yolox_forward_eval(images, sd)
    ref = T(g["eval_out"])
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-4), (out - ref).abs().max()
    det = orc.postprocess(out, 80, 0.001, 0.65)
    assert [0 if d is None else d.shape[0] for d in det] == list(g["eval_num_det"])


def test_preprocess_pads_with_114():
    ims = [torch.full((3, 50, 70), 7, dtype=torch.uint8), torch.full((3, 64, 40), 9, dtype=torch.uint8)]
    x = orc.preprocess(ims)
    assert x.shape == (2, 3, 64, 96) and x.dtype == torch.float32
    assert (x[0, :, :50, :70] == 7).all() and (x[0, :, 50:, :] == 114).all() and (x[1, :, :, 40:] == 114).all()


def test_state_dict_matches_reference_layout():
    """names / shapes of the seeded YOLOX-s state_dict are the reference's (8.97 M parameters)"""
    sd = orc.yolox_state_dict(0)
    n = sum(v.numel() for k, v in sd.items() if v.dtype == torch.float32 and "running" not in k)
    assert n == 8968255 or abs(n - 8.97e6) < 2e4, n
    g = load("model_w025.npz")
    ref_keys = {k[3:] for k in g.files if k.startswith("sd.")}
    assert set(orc.yolox_state_dict(0, width=0.25).keys()) == ref_keys
