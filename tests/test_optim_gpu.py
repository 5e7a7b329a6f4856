"""Flat SGD / AdamW step kernels against torch.optim (the arithmetic the reference's optimizer builders select,
yolov7/optimizer/build.py:234-256) on identical parameters, gradients and parameter groups.
Tolerance: 2e-6 relative + 1e-7 absolute per step (fp32; torch fuses a few multiply-adds differently)."""
import pytest
import torch

from yolov7_d2_b200 import optim

pytestmark = pytest.mark.gpu


def _setup(seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [("m0.conv.weight", (16, 3, 3, 3)), ("m0.bn.weight", (16,)), ("m0.bn.bias", (16,)), ("m1.conv.weight", (33, 16, 1, 1)),
              ("head.pred.weight", (5, 33, 1, 1)), ("head.pred.bias", (5,))]
    layout, off, tensors = [], 0, {}
    for name, shp in shapes:
        n = 1
        for d in shp:
            n *= d
        layout.append((name, off, n))
        tensors[name] = torch.randn(shp, generator=g)
        off += n
        off = (off + 3) // 4 * 4 + (4 if "m1" in name else 0)  # alignment gaps, one of them larger
    return layout, off, tensors, g


def _flat(layout, total, tensors, dev):
    flat = torch.zeros(total)
    for name, off, n in layout:
        flat[off:off + n] = tensors[name].reshape(-1)
    return flat.to(dev)


@pytest.mark.parametrize("kind,nesterov,clip", [("sgd", False, 0.0), ("sgd", True, 0.0), ("sgd", False, 0.5), ("adamw", False, 0.0), ("adamw", False, 0.5)])
def test_step_matches_torch(kind, nesterov, clip, cuda):
    layout, total, tensors, g = _setup()
    wd, wd_norm, wd_bias, bias_lr = 5e-4, 0.0, 1e-5, 2.0
    segs = optim.param_segments(layout, total, wd, wd_norm, wd_bias, bias_lr)
    flat_p = _flat(layout, total, tensors, cuda)
    flat_g = torch.zeros(total, device=cuda)
    lr = 0.05 if kind == "sgd" else 1e-2
    opt = optim.FlatOptimizer(flat_p, flat_g, segs, lr, kind, momentum=0.9 if kind == "sgd" else 0.0, nesterov=nesterov, clip_norm=clip, grad_scale=0.5)
    # torch reference with the groups the reference's builder would create
    ref_p = {n: torch.nn.Parameter(t.clone()) for n, t in tensors.items()}
    groups = []
    for n, p in ref_p.items():
        gw = wd_norm if ".bn." in n else (wd_bias if n.endswith(".bias") else wd)
        glr = lr * (bias_lr if n.endswith(".bias") else 1.0)
        groups.append({"params": [p], "weight_decay": gw, "lr": glr})
    ref = torch.optim.SGD(groups, lr, momentum=0.9, nesterov=nesterov) if kind == "sgd" else torch.optim.AdamW(groups, lr)
    for step in range(4):
        grads = {n: torch.randn(t.shape, generator=g) * (3.0 if step == 1 else 0.3) for n, t in tensors.items()}
        flat_g.copy_(_flat(layout, total, grads, cuda))
        pad_before = flat_p.clone()
        opt.step()
        for n, p in ref_p.items():
            p.grad = grads[n] * 0.5  # grad_scale
        if clip > 0:
            torch.nn.utils.clip_grad_norm_(list(ref_p.values()), clip)
        ref.step()
        got = flat_p.cpu()
        for n, off, cnt in layout:
            torch.testing.assert_close(got[off:off + cnt].view(ref_p[n].shape), ref_p[n].detach(), rtol=2e-6 * (step + 1), atol=1e-7 * (step + 1), msg=lambda m: f"{n} step {step}: {m}")
        # padding between tensors is never touched
        mask = torch.ones(total, dtype=torch.bool)
        for n, off, cnt in layout:
            mask[off:off + cnt] = False
        assert torch.equal(got[mask], pad_before.cpu()[mask])


def test_grad_norm_is_deterministic_and_exact(cuda):
    g = torch.randn(1_000_003, device=cuda)
    layout = [("w", 0, g.numel())]
    opt = optim.FlatOptimizer(torch.zeros_like(g), g, optim.param_segments(layout, g.numel(), 0.0), 0.0, "sgd", clip_norm=1.0)
    opt.step()
    a = opt.total_norm.clone()
    opt.step()
    assert torch.equal(a, opt.total_norm)
    ref = g.double().norm().item()
    assert abs(a.item() - ref) <= 1e-6 * ref


def test_model_optimizer_step_changes_weights_like_torch(cuda):
    """sgd(cfg, engine) on the real YOLOX-s layout: one step on random gradients equals torch.optim.SGD on the state_dict tensors"""
    from yolov7_d2_b200.engine import YoloxEngine

    class S:
        BASE_LR, MOMENTUM, NESTEROV, WEIGHT_DECAY, WEIGHT_DECAY_NORM, OPTIMIZER = 0.02, 0.9, True, 5e-4, 0.0, "SGD"

    class C:
        SOLVER = S()

    eng = YoloxEngine(1, 64, 64, device=cuda)
    eng.init_weights(0)
    opt = optim.build_optimizer_mapper(C(), eng)
    eng.flat_grad.normal_(generator=torch.Generator(device=cuda).manual_seed(1))
    before = {n: eng.params[n].detach().cpu().clone() for n in eng.param_names}
    grads = {n: eng.grads[n].detach().cpu().clone() for n in eng.param_names}
    opt.step()
    ref_p = {n: torch.nn.Parameter(t.clone()) for n, t in before.items()}
    groups = [{"params": [p], "weight_decay": 0.0 if ".bn." in n else 5e-4} for n, p in ref_p.items()]
    ref = torch.optim.SGD(groups, 0.02, momentum=0.9, nesterov=True)
    for n, p in ref_p.items():
        p.grad = grads[n]
    ref.step()
    for n in eng.param_names:
        torch.testing.assert_close(eng.params[n].cpu(), ref_p[n].detach(), rtol=2e-6, atol=1e-7, msg=lambda m: f"{n}: {m}")


@pytest.mark.parametrize("kind", ["sgd", "adamw"])
def test_state_dict_round_trip_resumes_identically(kind, cuda):
    """checkpoint / resume (detectron2's DetectionCheckpointer saves optimizer.state_dict()): momentum / moments / step count travel, so the
    step after a resume equals the step of the uninterrupted run"""
    layout, total, tensors, g = _setup(3)
    segs = optim.param_segments(layout, total, 5e-4, 0.0)
    grads = [_flat(layout, total, {n: torch.randn(t.shape, generator=g) for n, t in tensors.items()}, cuda) for _ in range(3)]

    def make():
        p = _flat(layout, total, tensors, cuda)
        gbuf = torch.zeros(total, device=cuda)
        return p, gbuf, optim.FlatOptimizer(p, gbuf, segs, 0.05 if kind == "sgd" else 1e-2, kind, momentum=0.9 if kind == "sgd" else 0.0)

    p1, g1, o1 = make()
    for k in range(3):
        g1.copy_(grads[k])
        o1.step()
    p2, g2, o2 = make()
    for k in range(2):
        g2.copy_(grads[k])
        o2.step()
    saved = o2.state_dict()
    assert saved["state"][0]["step"] == 2 and any(torch.is_tensor(v) and v.abs().sum() > 0 for v in saved["state"][0].values())
    p3, g3, o3 = make()
    p3.copy_(p2)
    o3.load_state_dict(saved)
    g3.copy_(grads[2])
    o3.step()
    assert o3.steps == 3
    assert torch.equal(p3, p1), "resumed step differs from the uninterrupted run"


def test_grad_scaler_drives_the_flat_optimizer(cuda):
    """torch.amp.GradScaler (detectron2 AMPTrainer, SOLVER.AMP.ENABLED) unscales / inf-checks through flat_param.grad = flat_grad:
    a scaled gradient gives the same update as the unscaled one, an inf gradient skips the step and halves the scale"""
    layout, total, tensors, g = _setup(4)
    segs = optim.param_segments(layout, total, 0.0)
    p = _flat(layout, total, tensors, cuda)
    gbuf = torch.zeros(total, device=cuda)
    opt = optim.FlatOptimizer(p, gbuf, segs, 0.1, "sgd", momentum=0.0)
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
    assert float(scaler.scale(torch.ones(1, device=cuda))) == 1024.0  # scale(loss) is what initialises the scaler's device state
    grad = torch.randn(total, device=cuda, generator=torch.Generator(device=cuda).manual_seed(5))
    p0 = p.clone()
    gbuf.copy_(grad * 1024.0)  # what backward of scaler.scale(loss) leaves in the flat buffer
    scaler.step(opt)
    scaler.update()
    mask = torch.zeros(total, dtype=torch.bool)
    for _, off, n in layout:
        mask[off:off + n] = True
    mask = mask.to(cuda)
    torch.testing.assert_close(p[mask], (p0 - 0.1 * grad)[mask], rtol=1e-6, atol=1e-7)
    p1 = p.clone()
    gbuf.copy_(grad * 1024.0)
    gbuf[3] = float("inf")
    scaler.step(opt)
    scaler.update()
    assert torch.equal(p, p1) and scaler.get_scale() == 512.0
