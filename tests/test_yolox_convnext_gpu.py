"""BASELINE.json configs[2]: YOLOX on a ConvNeXt-T backbone (corrected wiring: ConvNeXt stages 1-3 -> PAFPN / head width 0.75,
yolov7_d2_b200/yolox_convnext.py) against the composed CPU oracles (oracle/convnext_oracle.py -> oracle/yolox_oracle.py pafpn / head /
SimOTA / losses).  Yardstick as in tests/test_engine_gpu.py: the oracles with 16-bit storage emulation."""
import numpy as np
import pytest
import torch

from oracle import convnext_oracle as cno
from oracle import yolox_oracle as orc

pytestmark = pytest.mark.gpu


def _state(seed):
    sd = {k: v for k, v in orc.yolox_state_dict(seed, width=0.75).items() if not k.startswith("backbone.")}
    cn = cno.convnext_state_dict(seed + 1, trained_like=True)
    sd.update({"backbone." + k: v for k, v in cn.items()})
    return sd


def _oracle(sd, images, labels, emulate, g_raw):
    """head outputs and loss of the composed oracles; parameter gradients of the LINEAR functional sum(g_raw * raw head outputs) -- no discrete
    decision (SimOTA) between parameters and objective, as in tests/test_engine_gpu.py::test_backward_against_oracle"""
    orc.EMULATE_STORAGE = cno.EMULATE_STORAGE = emulate
    try:
        sd = {k: v.clone() for k, v in sd.items()}
        for k, v in sd.items():
            if v.dtype == torch.float32 and "running" not in k:
                v.requires_grad_(True)
        f1, f2, f3 = cno.forward_features(images.float(), sd, out_indices=(1, 2, 3), prefix="backbone.")
        raw = orc.head_raw(orc.pafpn({"dark3": f1, "dark4": f2, "dark5": f3}, sd, True), sd, True)
        outputs = orc.decode_train([r.detach() for r in raw])
        xs, ys, ss = orc.anchor_grid([o.shape[-2:] for o in raw])
        total = orc.yolox_losses(outputs, labels, xs, ys, ss)[0]
        flat = torch.cat([r.permute(0, 2, 3, 1).reshape(r.shape[0], -1, r.shape[1]) for r in raw], 1)
        (flat * g_raw).sum().backward()
    finally:
        orc.EMULATE_STORAGE = cno.EMULATE_STORAGE = False
    return outputs.detach(), {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}, float(total)


def test_yolox_convnext_step_against_oracles(cuda):
    from yolov7_d2_b200.yolox_convnext import YoloxConvNeXtEngine

    batch, size = 4, 256
    sd = _state(41)
    images, labels = orc.synthetic_batch(batch, size, 42, max_gt=6)
    eng = YoloxConvNeXtEngine(batch, size, size, device=cuda)
    eng.load_state_dict(sd)
    eng.images_u8.copy_(images.to(cuda))
    eng.labels.copy_(labels.to(cuda))
    eng.train_step()
    torch.cuda.synchronize()
    out = eng.outputs.cpu()
    n, a, ch = out.shape
    g_raw = (torch.randn(n, a, ch, generator=torch.Generator().manual_seed(43)) * 1e-2).to(torch.bfloat16).float()
    ref_out, ref_g, ref_loss = _oracle(sd, images, labels, False, g_raw)
    emu_out, emu_g, emu_loss = _oracle(sd, images, labels, True, g_raw)
    e_eng, e_emu = (out - ref_out).abs()[..., 4:].mean(), (emu_out - ref_out).abs()[..., 4:].mean()
    print("YOLOX-ConvNeXt logits vs fp32 oracles: engine mean err %.5f, 16-bit-storage oracles %.5f; loss %.4f / %.4f / %.4f" %
          (e_eng, e_emu, float(eng.losses[0]), ref_loss, emu_loss))
    assert e_eng <= 1.5 * e_emu + 1e-3
    # SimOTA + losses on the engine's own head outputs: bit-exact indices, 1e-4 losses
    xs, ys, ss = orc.anchor_grid([(h, w) for h, w, _, _ in eng.levels])
    total, iou5, lobj, lcls, ratio, assigns = orc.yolox_losses(out, labels, xs, ys, ss, return_assign=True)
    fg = eng.fg_mask.cpu().bool()
    for b, (rfg, mgt, _, _) in enumerate(assigns):
        assert torch.equal(fg[b], rfg) and torch.equal(eng.matched_gt.cpu()[b][rfg].long(), mgt)
    got = eng.losses.cpu().double().numpy()
    assert np.allclose(got[:4], [float(total), float(iou5), float(lobj), float(lcls)], rtol=1e-4, atol=1e-5)
    # backward of both plans for the fixed upstream gradient g_raw on the raw head outputs
    yx = eng.yx
    eng.pack_weights()
    eng.forward_features(True)
    for k, (h, w, s_, a_off) in enumerate(yx.levels):
        gl = g_raw[:, a_off:a_off + h * w]
        yx.d_cls[k].copy_(gl[..., 5:].reshape(n, h, w, ch - 5).to(cuda))
        yx.d_ro[k].zero_()
        yx.d_ro[k][..., :5].copy_(gl[..., :5].reshape(n, h, w, 5).to(cuda))
        yx.bias_acc[k].copy_(gl.double().sum((0, 1)).to(cuda))
    eng.backward()
    torch.cuda.synchronize()
    worst = []
    for name in eng.param_names:
        if name not in ref_g:
            continue
        g = eng.grads[name].cpu().flatten().double()
        r, e = ref_g[name].flatten().double(), emu_g[name].flatten().double()
        if r.norm() == 0:
            continue
        cos = float((g @ r) / (g.norm() * r.norm() + 1e-30))
        cos_e = float((e @ r) / (e.norm() * r.norm() + 1e-30))
        worst.append((cos - cos_e, cos, cos_e, name))
    worst.sort()
    print("largest cosine deficit vs the 16-bit-storage oracles:", worst[:5])
    assert len(worst) > 150
    # a gradient that the 16-bit-storage ORACLE itself cannot reproduce (sums of many cancelling terms, e.g. LayerNorm biases: cosine of the
    # emulating oracle to the fp32 oracle near 0) carries no signal to compare; judge the well-conditioned ones
    conditioned = [(cos, cos_e, name) for _, cos, cos_e, name in worst if cos_e >= 0.8]
    assert len(conditioned) >= 0.6 * len(worst), (len(conditioned), len(worst))
    for cos, cos_e, name in conditioned:
        assert cos >= cos_e - 0.1, (name, cos, cos_e)


def test_yolox_meta_arch_with_convnext_backbone(cuda):
    """`MODEL.BACKBONE.NAME: build_convnext_backbone` + `META_ARCHITECTURE: YOLOX` (configs/coco/yolox/yolox_convnext.yaml) builds and trains"""
    import bench
    from yolov7_d2_b200.modeling import YOLOX
    from yolov7_d2_b200 import optim

    cfg = bench.yolox_s_cfg("cuda")
    cfg.MODEL.BACKBONE.NAME = "build_convnext_backbone"
    m = YOLOX(cfg)
    m.train()
    names = dict(m.named_parameters())
    assert "backbone.stages.2.8.pwconv1.weight" in names and "neck.C3_p4.conv1.conv.weight" in names and names["head.stems.0.conv.weight"].shape[0] == 192
    images, labels = orc.synthetic_batch(2, 128, 51, max_gt=4)
    opt = optim.build_optimizer_mapper(cfg, m)
    before = names["backbone.stages.1.0.pwconv1.weight"].detach().clone()
    out = m(bench.batched_inputs_from(images, labels))
    sum(out.values()).backward()
    opt.step()
    assert torch.isfinite(out["total_loss"]) and not torch.equal(before, names["backbone.stages.1.0.pwconv1.weight"].detach())
    m.eval()
    res = m(bench.batched_inputs_from(images, labels))
    assert len(res) == 2 and hasattr(res[0]["instances"], "pred_boxes")
