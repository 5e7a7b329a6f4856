"""GPU parity of decode / SimOTA / loss kernels (C ABI) against the CPU oracle and the reference-generated golden
fixtures.  Index outputs (fg_mask, matched gt / class) must be bit-exact; IoUs, losses and gradients are fp32."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import yolox_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_assign_and_loss(capi, outputs, labels, hw_strides, dev, weights=(5.0, 1.0, 1.0)):
    """returns dict of torch CPU tensors"""
    L = capi.lib()
    b, a, ch = outputs.shape
    out = outputs.to(dev).contiguous()
    lab = labels.to(dev).contiguous()
    lv = (ctypes.c_int32 * (3 * len(hw_strides)))(*[v for t in hw_strides for v in t])
    ws = torch.empty(L.yb200_simota_workspace(b, a), dtype=torch.uint8, device=dev)
    num_gt = torch.empty(b, dtype=torch.int32, device=dev)
    fg = torch.empty(b, a, dtype=torch.uint8, device=dev)
    mgt = torch.empty(b, a, dtype=torch.int32, device=dev)
    miou = torch.empty(b, a, dtype=torch.float32, device=dev)
    mcls = torch.empty(b, a, dtype=torch.int32, device=dev)
    nfg = torch.empty(b, dtype=torch.int32, device=dev)
    totals = torch.empty(2, dtype=torch.int32, device=dev)
    capi.check(L.yb200_simota_assign(capi.ptr(out), capi.ptr(lab), b, a, ch, lab.shape[1], lv, len(hw_strides), capi.ptr(ws), capi.ptr(num_gt),
                                     capi.ptr(fg), capi.ptr(mgt), capi.ptr(miou), capi.ptr(mcls), capi.ptr(nfg), capi.ptr(totals),
                                     capi.stream_ptr()), "simota_assign")
    w3 = torch.tensor(weights, dtype=torch.float32, device=dev)
    acc = torch.zeros(3, dtype=torch.float64, device=dev)
    losses = torch.empty(6, dtype=torch.float32, device=dev)
    dense = torch.full((b, a, ch), float("nan"), dtype=torch.float32, device=dev)
    nc = ch - 5
    d_cls = [torch.full((b, h, w, nc), float("nan"), dtype=torch.bfloat16, device=dev) for h, w, _ in hw_strides]
    d_ro = [torch.full((b, h, w, 16), float("nan"), dtype=torch.bfloat16, device=dev) for h, w, _ in hw_strides]
    pc = (ctypes.c_void_p * len(hw_strides))(*[t.data_ptr() for t in d_cls])
    pr = (ctypes.c_void_p * len(hw_strides))(*[t.data_ptr() for t in d_ro])
    bias_acc = torch.zeros(len(hw_strides), ch, dtype=torch.float64, device=dev)
    capi.check(L.yb200_yolox_loss(capi.ptr(out), capi.ptr(lab), b, a, ch, lab.shape[1], lv, len(hw_strides), capi.ptr(fg), capi.ptr(mgt),
                                  capi.ptr(miou), capi.ptr(mcls), capi.ptr(totals), capi.ptr(w3), capi.ptr(acc), capi.ptr(losses), pc, pr,
                                  capi.ptr(dense), capi.ptr(bias_acc), capi.stream_ptr()), "yolox_loss")
    torch.cuda.synchronize()
    return dict(num_gt=num_gt.cpu(), fg=fg.cpu().bool(), mgt=mgt.cpu(), miou=miou.cpu(), mcls=mcls.cpu(), nfg=nfg.cpu(), totals=totals.cpu(),
                losses=losses.cpu(), dense=dense.cpu(), d_cls=[t.cpu() for t in d_cls], d_ro=[t.cpu() for t in d_ro], bias=bias_acc.cpu(),
                acc=acc.cpu())


def raw_grad_from_decoded(out, grad_dec, hw_strides):
    """chain rule through the decode: d/d raw_xy = d/d xy * s, d/d raw_wh = d/d wh * wh"""
    g = grad_dec.clone()
    off = 0
    for h, w, s in hw_strides:
        g[:, off:off + h * w, :2] *= s
        off += h * w
    g[..., 2:4] = grad_dec[..., 2:4] * out[..., 2:4]
    return g


@pytest.mark.parametrize("case", ["trained", "init", "crowd"])
def test_simota_loss_vs_reference_golden(cuda, case):
    from yolov7_d2_b200 import capi

    g = np.load(os.path.join(GOLD, "simota.npz"))
    size = int(g["size"])
    hw = [(size // s, size // s, s) for s in orc.STRIDES]
    out = torch.from_numpy(g[f"{case}.outputs"])
    labels = torch.from_numpy(g[f"{case}.labels"])
    r = run_assign_and_loss(capi, out, labels, hw, cuda)
    with_gt = set(int(b) for b in g[f"{case}.images_with_gt"])
    for b in range(out.shape[0]):
        if b not in with_gt:
            assert not r["fg"][b].any() and int(r["num_gt"][b]) == 0
            continue
        ref_fg = torch.from_numpy(g[f"{case}.b{b}.fg_mask"])
        assert torch.equal(r["fg"][b], ref_fg), f"fg_mask image {b}: {(r['fg'][b] != ref_fg).sum()} differ"
        sel = ref_fg
        assert torch.equal(r["mgt"][b][sel].long(), torch.from_numpy(g[f"{case}.b{b}.matched_gt"])), f"matched_gt image {b}"
        assert torch.equal(r["mcls"][b][sel].float(), torch.from_numpy(g[f"{case}.b{b}.matched_cls"])), f"matched_cls image {b}"
        assert torch.equal(r["miou"][b][sel], torch.from_numpy(g[f"{case}.b{b}.matched_iou"])), f"matched_iou image {b}"
        assert (r["mgt"][b][~sel] == -1).all()
    ref_l = g[f"{case}.losses"]
    got = r["losses"].double().numpy()
    assert np.allclose(got[[0, 1, 2, 3, 5]], ref_l, rtol=1e-4, atol=1e-5), (got, ref_l)   # north_star: losses within 1e-3 relative
    ref_grad = raw_grad_from_decoded(out, torch.from_numpy(g[f"{case}.grad"]), hw)
    err = (r["dense"] - ref_grad).abs().max().item()
    assert err <= 1e-4 * ref_grad.abs().max().item() + 1e-7, err
    # bf16 per-level gradient tensors and bias sums agree with the dense gradient
    off = 0
    for l, (h, w, s) in enumerate(hw):
        d = r["dense"][:, off:off + h * w]
        assert torch.allclose(r["d_cls"][l].float().reshape(d.shape[0], h * w, -1), d[..., 5:], rtol=2 ** -7, atol=1e-9)
        ro = r["d_ro"][l].float().reshape(d.shape[0], h * w, 16)
        assert torch.allclose(ro[..., :5], d[..., :5], rtol=2 ** -7, atol=1e-9) and (ro[..., 5:] == 0).all()
        assert torch.allclose(r["bias"][l].float(), d.sum((0, 1)), rtol=1e-4, atol=1e-6)
        off += h * w
    assert (r["acc"] == 0).all()


@pytest.mark.parametrize("case", ["trained", "init"])
def test_l1_branch_vs_reference_golden(cuda, case):
    """`use_l1` (yolox_head.py:186-195, 389-429, 443-448): yb200_yolox_decode_keep_raw + yb200_yolox_loss_l1 on the RAW head outputs of the fixture
    tests/golden/simota_l1.npz (reference head with origin_preds): the six losses to 1e-4, the gradient w.r.t. the raw outputs to 1e-4 of its max"""
    from yolov7_d2_b200 import capi

    L = capi.lib()
    g = np.load(os.path.join(GOLD, "simota_l1.npz"))
    size = int(g["size"])
    hw = [(size // s, size // s, s) for s in orc.STRIDES]
    raw = torch.from_numpy(g[f"{case}.raw"])
    labels = torch.from_numpy(g[f"{case}.labels"]).to(cuda).contiguous()
    b, a, ch = raw.shape
    out = raw.to(cuda).contiguous()
    lv = (ctypes.c_int32 * (3 * len(hw)))(*[v for t in hw for v in t])
    raw_reg = torch.full((b, a, 4), float("nan"), device=cuda)
    capi.check(L.yb200_yolox_decode_keep_raw(capi.ptr(out), b, a, ch, lv, len(hw), capi.ptr(raw_reg), capi.stream_ptr()), "decode_keep_raw")
    assert torch.equal(raw_reg.cpu(), raw[..., :4])
    ws = torch.empty(L.yb200_simota_workspace(b, a), dtype=torch.uint8, device=cuda)
    i32 = lambda *shape: torch.empty(*shape, dtype=torch.int32, device=cuda)
    num_gt, mgt, mcls, nfg, totals = i32(b), i32(b, a), i32(b, a), i32(b), i32(2)
    fg = torch.empty(b, a, dtype=torch.uint8, device=cuda)
    miou = torch.empty(b, a, device=cuda)
    capi.check(L.yb200_simota_assign(capi.ptr(out), capi.ptr(labels), b, a, ch, labels.shape[1], lv, len(hw), capi.ptr(ws), capi.ptr(num_gt), capi.ptr(fg),
                                     capi.ptr(mgt), capi.ptr(miou), capi.ptr(mcls), capi.ptr(nfg), capi.ptr(totals), capi.stream_ptr()), "simota_assign")
    w4 = torch.tensor([5.0, 1.0, 1.0, 1.0], device=cuda)
    acc = torch.zeros(4, dtype=torch.float64, device=cuda)
    losses = torch.empty(6, device=cuda)
    dense = torch.full((b, a, ch), float("nan"), device=cuda)
    capi.check(L.yb200_yolox_loss_l1(capi.ptr(out), capi.ptr(raw_reg), capi.ptr(labels), b, a, ch, labels.shape[1], lv, len(hw), capi.ptr(fg), capi.ptr(mgt),
                                     capi.ptr(miou), capi.ptr(mcls), capi.ptr(totals), capi.ptr(w4), capi.ptr(acc), capi.ptr(losses), None, None,
                                     capi.ptr(dense), None, capi.stream_ptr()), "yolox_loss_l1")
    torch.cuda.synchronize()
    got, ref = losses.cpu().double().numpy(), g[f"{case}.losses"]
    assert np.allclose(got, ref, rtol=1e-4, atol=1e-5), (got, ref)
    assert got[4] > 0.1 and abs(got[0] - (got[1] + got[2] + got[3] + got[4])) <= 1e-4 * got[0]
    ref_grad = torch.from_numpy(g[f"{case}.grad"])
    err = (dense.cpu() - ref_grad).abs().max().item()
    assert err <= 1e-4 * ref_grad.abs().max().item() + 1e-7, err
    assert (acc.cpu() == 0).all()


def test_simota_full_size_vs_oracle(cuda):
    """640x640 (8400 anchors), batch 4 incl. an empty image and a crowded one: exact indices vs the CPU oracle"""
    from oracle.gen_golden import trained_like_outputs
    from yolov7_d2_b200 import capi

    size = 640
    hw = [(size // s, size // s, s) for s in orc.STRIDES]
    _, labels = orc.synthetic_batch(4, size, 77, max_gt=60, empty_every=4)
    out = trained_like_outputs(None, labels, size, 80, 78)
    xs, ys, ss = orc.anchor_grid([(h, w) for h, w, _ in hw])
    o = out.clone().requires_grad_(True)
    total, iou5, lobj, lcls, ratio, assigns = orc.yolox_losses(o, labels, xs, ys, ss, return_assign=True)
    total.backward()
    r = run_assign_and_loss(capi, out, labels, hw, cuda)
    nfg = 0
    for b, (fg, mgt, mcls, miou) in enumerate(assigns):
        assert torch.equal(r["fg"][b], fg), f"fg image {b}: {(r['fg'][b] != fg).sum().item()} differ"
        assert torch.equal(r["mgt"][b][fg].long(), mgt) and torch.equal(r["mcls"][b][fg].float(), mcls)
        assert torch.equal(r["miou"][b][fg], miou)
        assert int(r["nfg"][b]) == int(fg.sum())
        nfg += int(fg.sum())
    assert int(r["totals"][0]) == nfg and int(r["totals"][1]) == int((labels.sum(2) > 0).sum())
    got = r["losses"].double().numpy()
    ref = np.array([float(total), float(iou5), float(lobj), float(lcls), 0.0, float(ratio)])
    assert np.allclose(got, ref, rtol=1e-4, atol=1e-5), (got, ref)
    ref_grad = raw_grad_from_decoded(out, o.grad, hw)
    assert (r["dense"] - ref_grad).abs().max().item() <= 1e-4 * ref_grad.abs().max().item() + 1e-7


@pytest.mark.parametrize("eval_mode", [0, 1])
def test_decode(cuda, eval_mode):
    from yolov7_d2_b200 import capi

    g = torch.Generator().manual_seed(5)
    hw = [(8, 12, 8), (4, 6, 16), (2, 3, 32)]
    raw = [torch.randn(3, 85, h, w, generator=g) for h, w, _ in hw]
    ref = (orc.decode_eval if eval_mode else orc.decode_train)(raw, [s for _, _, s in hw])
    flat = torch.cat([r.permute(0, 2, 3, 1).reshape(3, -1, 85) for r in raw], 1).contiguous().to(cuda)
    lv = (ctypes.c_int32 * 9)(*[v for t in hw for v in t])
    capi.check(capi.lib().yb200_yolox_decode(capi.ptr(flat), 3, flat.shape[1], 85, lv, 3, eval_mode, capi.stream_ptr()), "decode")
    assert torch.allclose(flat.cpu(), ref, rtol=1e-6, atol=1e-6)
