"""GPU parity AT THE HEADLINE CONFIGURATION (BASELINE.json configs[1]: YOLOX-s, 640x640): the plan that bench.py times -- full-size
kernels, 51 200-tile persistent walks, the stem weight gradient over 6.5 M pixels -- against the CPU oracle.

  * 8 x 640x640: forward head logits (16-bit-storage yardstick of tests/test_engine_gpu.py), SimOTA + losses on the engine's own head
    outputs (indices bit-exact, losses 1e-4), parameter gradients of a linear functional vs the fp32 oracle;
  * 64 x 640x640 (the benchmark batch): one training step; SimOTA indices bit-exact and losses 1e-4 against the oracle evaluated on
    the engine's head outputs; forward logits against the fp32 oracle forward pass; finite, reproducible gradients.
The strict (split-bf16) forward is checked against the literal 1e-3 of north_star in tests/test_strict_gpu.py.
"""
import numpy as np
import pytest
import torch

from oracle import yolox_oracle as orc

pytestmark = pytest.mark.gpu


def _sd(seed):
    sd = orc.yolox_state_dict(seed)
    g = torch.Generator().manual_seed(seed + 50)
    for k in sd:
        if k.endswith(".bn.weight"):
            sd[k] = torch.rand(sd[k].shape, generator=g) * 0.5 + 0.75
        if k.endswith(".bn.bias"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
    return sd


def _check_assign_and_losses(eng, labels):
    out = eng.outputs.cpu()
    xs, ys, ss = orc.anchor_grid([(h, w) for h, w, _, _ in eng.levels])
    total, iou5, lobj, lcls, ratio, assigns = orc.yolox_losses(out, labels, xs, ys, ss, return_assign=True)
    fg = eng.fg_mask.cpu().bool()
    mg, mi, mc = eng.matched_gt.cpu(), eng.matched_iou.cpu(), eng.matched_cls.cpu()
    nfg = 0
    for b, (rfg, mgt, mcls, miou) in enumerate(assigns):
        assert torch.equal(fg[b], rfg), f"image {b}: foreground mask"
        assert torch.equal(mg[b][rfg].long(), mgt), f"image {b}: matched gt"
        assert torch.equal(mc[b][rfg].long(), mcls.long()), f"image {b}: matched class"
        assert torch.equal(mi[b][rfg], miou), f"image {b}: matched IoU"
        nfg += int(rfg.sum())
    got = eng.losses.cpu().double().numpy()
    ref = np.array([float(total), float(iou5), float(lobj), float(lcls), 0.0, float(ratio)])
    assert np.allclose(got, ref, rtol=1e-4, atol=1e-5), (got, ref)
    return nfg


@pytest.fixture(scope="module")
def step8(cuda):
    from yolov7_d2_b200.engine import YoloxEngine

    sd = _sd(11)
    images, labels = orc.synthetic_batch(8, 640, 21, max_gt=12, empty_every=8)
    eng = YoloxEngine(8, 640, 640, device=cuda)
    eng.load_state_dict(sd)
    eng.images_u8.copy_(images.to(cuda))
    eng.labels.copy_(labels.to(cuda))
    eng.train_step()
    torch.cuda.synchronize()
    return dict(eng=eng, sd=sd, images=images, labels=labels)


def _fwd(sd, images, labels, emulate):
    orc.EMULATE_STORAGE = emulate
    try:
        with torch.no_grad():
            return orc.yolox_forward_train(images.float(), labels, {k: v.clone() for k, v in sd.items()})[-1]
    finally:
        orc.EMULATE_STORAGE = False


def test_640_bs8_forward_logits(step8):
    eng = step8["eng"]
    ref = _fwd(step8["sd"], step8["images"], step8["labels"], False)
    emu = _fwd(step8["sd"], step8["images"], step8["labels"], True)
    out = eng.outputs.cpu()
    e_eng, e_emu = (out - ref).abs()[..., 4:], (emu - ref).abs()[..., 4:]
    print("640x640 bs8 logit error vs fp32 oracle: engine mean %.5f max %.4f | 16-bit-storage oracle mean %.5f max %.4f" %
          (e_eng.mean(), e_eng.max(), e_emu.mean(), e_emu.max()))
    assert e_eng.mean() <= 1.5 * e_emu.mean() + 1e-4
    a, b = out[..., 4:].flatten().double(), ref[..., 4:].flatten().double()
    assert float(((a - a.mean()) * (b - b.mean())).mean() / (a.std() * b.std())) >= 0.99


def test_640_bs8_simota_and_losses(step8):
    assert _check_assign_and_losses(step8["eng"], step8["labels"]) > 0


def test_640_bs8_backward_vs_oracle(step8, cuda):
    """whole-network backward at 640x640 for a fixed upstream gradient on the raw head outputs (same functional as
    tests/test_engine_gpu.py::test_backward_against_oracle, real layer shapes)"""
    eng, images, sd0 = step8["eng"], step8["images"], step8["sd"]
    gen = torch.Generator().manual_seed(78)
    n, a, ch = eng.outputs.shape
    g_raw = (torch.randn(n, a, ch, generator=gen) * 1e-2).to(torch.bfloat16).float()
    eng.load_state_dict(sd0)
    eng.pack_weights()
    eng.preprocess()
    eng.forward_features(True)
    for k, (h, w, s, a_off) in enumerate(eng.levels):
        gl = g_raw[:, a_off:a_off + h * w]
        eng.d_cls[k].copy_(gl[..., 5:].reshape(n, h, w, ch - 5).to(cuda))
        eng.d_ro[k].zero_()
        eng.d_ro[k][..., :5].copy_(gl[..., :5].reshape(n, h, w, 5).to(cuda))
        eng.bias_acc[k].copy_(gl.double().sum((0, 1)).to(cuda))
    eng.backward()
    torch.cuda.synchronize()
    def oracle_grads(emulate):
        orc.EMULATE_STORAGE = emulate
        try:
            sd = {k: v.clone() for k, v in sd0.items()}
            for k, v in sd.items():
                if v.dtype == torch.float32 and "running" not in k:
                    v.requires_grad_(True)
            raw = orc.head_raw(orc.pafpn(orc.csp_darknet(images.float(), sd, True), sd, True), sd, True)
            flat = torch.cat([r.permute(0, 2, 3, 1).reshape(r.shape[0], -1, r.shape[1]) for r in raw], 1)
            (flat * g_raw).sum().backward()
        finally:
            orc.EMULATE_STORAGE = False
        return {k: v.grad for k, v in sd.items() if v.requires_grad}

    ref, emu = oracle_grads(False), oracle_grads(True)  # fp32 oracle; oracle rounding at the engine's 16-bit storage points (the yardstick)
    worst = []
    for name in eng.param_names:
        g = eng.grads[name].cpu().flatten().double()
        r, e = ref[name].flatten().double(), emu[name].flatten().double()
        cos = float((g @ r) / (g.norm() * r.norm() + 1e-30))
        cos_e = float((e @ r) / (e.norm() * r.norm() + 1e-30))
        worst.append((cos, cos_e, float(g.norm() / (r.norm() + 1e-30)), name))
    worst.sort()
    print("lowest cosine similarity to the fp32 oracle at 640x640 (engine, 16-bit-storage oracle, norm ratio):", worst[:6])
    for cos, cos_e, ratio, name in worst:
        assert cos >= min(0.9, cos_e - 0.05) and cos >= 0.75 and 0.8 <= ratio <= 1.25, (name, cos, cos_e, ratio)


def test_640_bs64_benchmark_step(cuda):
    """the exact plan bench.py replays: 64 x 640x640"""
    from yolov7_d2_b200.engine import YoloxEngine

    sd = _sd(12)
    images, labels = orc.synthetic_batch(64, 640, 22, max_gt=20, empty_every=16)
    eng = YoloxEngine(64, 640, 640, device=cuda)
    eng.load_state_dict(sd)
    eng.images_u8.copy_(images.to(cuda))
    eng.labels.copy_(labels.to(cuda))
    eng.train_step()
    torch.cuda.synchronize()
    nfg = _check_assign_and_losses(eng, labels)
    assert nfg == int(eng.totals[0]) and nfg > 64
    g1, l1 = eng.flat_grad.clone(), eng.losses.clone()
    assert torch.isfinite(g1).all() and float(g1.abs().sum()) > 0
    # forward logits against the fp32 oracle forward pass of the same 64 images (16-bit storage bound of DESIGN.md par.2)
    ref = _fwd(sd, images, labels, False)
    err = (eng.outputs.cpu() - ref).abs()[..., 4:]
    print("640x640 bs64 logit error vs fp32 oracle: mean %.5f max %.4f" % (err.mean(), err.max()))
    # 16-bit storage through ~60 layers (tests/test_engine_gpu.py measures the same network at 0.025 with its yardstick 0.024); the maximum over
    # 43 M logits is an outlier statistic (2.5 observed), the 99.99th percentile is the robust bound
    q = torch.quantile(err.flatten()[:: 37].float(), 0.9999).item()
    print("99.99th percentile %.4f" % q)
    assert err.mean() <= 0.035 and q <= 1.2
    eng.load_state_dict(sd)
    eng.train_step()
    torch.cuda.synchronize()
    assert torch.allclose(l1, eng.losses, rtol=1e-6)
    assert float((g1 - eng.flat_grad).abs().max() / g1.abs().max()) <= 2e-3
    del eng
    torch.cuda.empty_cache()
