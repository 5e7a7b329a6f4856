"""Layer-by-layer comparison of the engine's activations with the CPU oracle (fp32 and bf16-rounded variants)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import yolox_oracle as orc
from yolov7_d2_b200.engine import YoloxEngine, ConvOp

dev = torch.device("cuda:0")
batch, size = 4, 128
sd = orc.yolox_state_dict(3)
g = torch.Generator().manual_seed(9)
for k in sd:
    if k.endswith(".bn.weight"): sd[k] = torch.rand(sd[k].shape, generator=g) * 0.5 + 0.75
    if k.endswith(".bn.bias"): sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
images, labels = orc.synthetic_batch(batch, size, 5, max_gt=6, empty_every=4)
eng = YoloxEngine(batch, size, size, device=dev)
eng.load_state_dict(sd)
eng.images_u8.copy_(images.to(dev)); eng.labels.copy_(labels.to(dev))
eng.train_step(); torch.cuda.synchronize()
for mode in (True, False):
    orc.EMULATE_STORAGE = mode
    orc.TRACE = {}
    ref_sd = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        res = orc.yolox_forward_train(images.float(), labels, ref_sd)
    print("=== oracle EMULATE_STORAGE =", mode, "losses", [float(x) for x in res[:4]], "engine", eng.losses.tolist())
    for op in eng.ops:
        if not isinstance(op, ConvOp): continue
        for hd in op.heads:
            a = hd.out.tensor().float().cpu()
            r = orc.TRACE[hd.prefix].permute(0, 2, 3, 1)
            err = (a - r).abs()
            print("%-34s max|ref| %7.3f  max err %8.4f  mean err %9.5f" % (hd.prefix, r.abs().max(), err.max(), err.mean()))
    out = eng.outputs.cpu()
    e = (out - res[-1]).abs()
    print("outputs: logits max err %.4f mean %.5f ; boxes max rel %.4f" % (e[..., 4:].max(), e[..., 4:].mean(), (e[..., :4] / res[-1][..., :4].abs().clamp(min=1)).max()))
