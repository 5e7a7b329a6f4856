"""End-to-end GPU parity of the YOLOX-s engine (forward, SimOTA + loss, backward) against the CPU oracle.

Tolerances (documented in DESIGN.md): activations are stored in bf16 (rel 2^-8 per layer, ~60 layers deep), so
  * head logits: max |err| <= 0.06 and mean |err| <= 0.01 against the fp32 oracle;
  * loss / SimOTA given the SAME head outputs: indices bit-exact, losses 1e-4 relative (fp32 kernels);
  * parameter gradients: cosine similarity >= 0.97 and norm ratio within 10 % per tensor (>= 0.995 for the head).
"""
import numpy as np
import pytest
import torch

from oracle import yolox_oracle as orc

pytestmark = pytest.mark.gpu


def _oracle_step(sd, images, labels):
    for k, v in sd.items():
        if v.dtype == torch.float32 and "running" not in k:
            v.requires_grad_(True)
    total, iou5, lobj, lcls, ratio, outputs = orc.yolox_forward_train(images.float(), labels, sd)
    total.backward()
    return dict(losses=np.array([float(total), float(iou5), float(lobj), float(lcls), float(ratio)]), outputs=outputs.detach())


@pytest.fixture(scope="module")
def step(cuda):
    from yolov7_d2_b200.engine import YoloxEngine

    torch.manual_seed(0)
    batch, size = 4, 128
    sd = orc.yolox_state_dict(3)
    # non-trivial BN affine parameters so that gamma/beta gradients and the scale/shift path are exercised
    g = torch.Generator().manual_seed(9)
    for k in sd:
        if k.endswith(".bn.weight"):
            sd[k] = (torch.rand(sd[k].shape, generator=g) * 0.5 + 0.75)
        if k.endswith(".bn.bias"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
    images, labels = orc.synthetic_batch(batch, size, 5, max_gt=6, empty_every=4)
    eng = YoloxEngine(batch, size, size, device=cuda)
    eng.load_state_dict(sd)
    eng.images_u8.copy_(images.to(cuda))
    eng.labels.copy_(labels.to(cuda))
    eng.train_step()
    torch.cuda.synchronize()
    ref_sd = {k: v.clone() for k, v in sd.items()}
    ref = _oracle_step(ref_sd, images, labels)
    return dict(eng=eng, ref=ref, ref_sd=ref_sd, images=images, labels=labels, sd0=sd)


def test_forward_logits(step):
    eng, ref = step["eng"], step["ref"]
    out = eng.outputs.cpu()
    err = (out - ref["outputs"]).abs()
    logits = err[..., 4:]
    assert logits.max() <= 0.06 and logits.mean() <= 0.01, (logits.max().item(), logits.mean().item())
    box_rel = err[..., :4] / ref["outputs"][..., :4].abs().clamp(min=1.0)
    assert box_rel.max() <= 0.05, box_rel.max().item()


def test_loss_and_simota_on_engine_outputs(step):
    """same inputs -> bit-exact assignment and fp32 losses (the reference's parity rule for SimOTA / losses)"""
    eng, labels = step["eng"], step["labels"]
    out = eng.outputs.cpu()
    xs, ys, ss = orc.anchor_grid([(h, w) for h, w, _, _ in eng.levels])
    total, iou5, lobj, lcls, ratio, assigns = orc.yolox_losses(out, labels, xs, ys, ss, return_assign=True)
    fg = eng.fg_mask.cpu().bool()
    for b, (rfg, mgt, mcls, miou) in enumerate(assigns):
        assert torch.equal(fg[b], rfg), f"image {b}"
        assert torch.equal(eng.matched_gt.cpu()[b][rfg].long(), mgt)
        assert torch.equal(eng.matched_iou.cpu()[b][rfg], miou)
    got = eng.losses.cpu().double().numpy()
    ref = np.array([float(total), float(iou5), float(lobj), float(lcls), 0.0, float(ratio)])
    assert np.allclose(got, ref, rtol=1e-4, atol=1e-5), (got, ref)


def test_losses_close_to_fp32_model(step):
    got = step["eng"].losses.cpu().double().numpy()[[0, 1, 2, 3, 5]]
    ref = step["ref"]["losses"]
    assert np.allclose(got, ref, rtol=3e-2, atol=1e-2), (got, ref)


def test_bn_running_stats(step):
    eng, ref_sd = step["eng"], step["ref_sd"]
    for name in ("backbone.stem.conv.bn", "backbone.dark3.1.m.1.conv2.bn", "neck.C3_n4.conv3.bn", "head.reg_convs.2.1.bn"):
        rm, rv = eng.buffers[name + ".running_mean"].cpu(), eng.buffers[name + ".running_var"].cpu()
        assert torch.allclose(rm, ref_sd[name + ".running_mean"], rtol=2e-2, atol=2e-3), name
        assert torch.allclose(rv, ref_sd[name + ".running_var"], rtol=2e-2, atol=2e-3), name
        assert int(eng.buffers[name + ".num_batches_tracked"]) == 1


def test_parameter_gradients(step):
    eng, ref_sd = step["eng"], step["ref_sd"]
    worst = []
    for name in eng.param_names:
        g = eng.grads[name].cpu().flatten().double()
        r = ref_sd[name].grad.flatten().double()
        cos = float((g @ r) / (g.norm() * r.norm() + 1e-30))
        ratio = float(g.norm() / (r.norm() + 1e-30))
        worst.append((cos, ratio, name))
        lim = 0.995 if name.startswith("head.") and "preds" in name else 0.97
        assert cos >= lim and 0.9 <= ratio <= 1.1, (name, cos, ratio)
    worst.sort()
    print("lowest cosine similarities:", worst[:5])


def test_second_step_accumulates_nothing_stale(step):
    """a second identical step reproduces the same gradients bit for bit (no stale accumulators, deterministic kernels)"""
    eng = step["eng"]
    eng.load_state_dict(step["sd0"])
    eng.train_step()
    torch.cuda.synchronize()
    g1 = eng.flat_grad.clone()
    l1 = eng.losses.clone()
    eng.load_state_dict(step["sd0"])
    eng.train_step()
    torch.cuda.synchronize()
    assert torch.equal(l1, eng.losses)
    rel = (g1 - eng.flat_grad).abs().max() / g1.abs().max()
    assert rel <= 1e-5, rel.item()   # fp64 atomics in the BN reductions may reorder; everything else is deterministic


def test_eval_forward_matches_oracle(step):
    eng, images = step["eng"], step["images"]
    sd = step["ref_sd"]
    eng.load_state_dict({k: v.detach() for k, v in sd.items()})
    out = eng.eval_forward().cpu()
    with torch.no_grad():
        ref = orc.yolox_forward_eval(images.float(), {k: v.detach() for k, v in sd.items()})
    err = (out - ref).abs()
    assert err[..., 4:].max() <= 0.02, err[..., 4:].max().item()   # probabilities
    assert (err[..., :4] / ref[..., :4].abs().clamp(min=1.0)).max() <= 0.05
