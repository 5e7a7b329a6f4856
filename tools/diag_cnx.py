"""diagnostic: YOLOX-ConvNeXt composite, per-layer activation error of the neck / head against the oracles (fp32 and 16-bit-storage)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import convnext_oracle as cno, yolox_oracle as orc
import test_yolox_convnext_gpu as T
from yolov7_d2_b200.yolox_convnext import YoloxConvNeXtEngine
from yolov7_d2_b200.engine import ConvOp

dev = torch.device("cuda:0")
torch.set_num_threads(16)
sd = T._state(41)
images, labels = orc.synthetic_batch(4, 256, 42, max_gt=6)
eng = YoloxConvNeXtEngine(4, 256, 256, device=dev)
eng.load_state_dict(sd)
eng.images_u8.copy_(images.to(dev)); eng.labels.copy_(labels.to(dev))
eng.train_step(); torch.cuda.synchronize()
def trace(emulate):
    orc.EMULATE_STORAGE = cno.EMULATE_STORAGE = emulate
    orc.TRACE = {}
    s2 = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        f = cno.forward_features(images.float(), s2, out_indices=(1, 2, 3), prefix="backbone.")
        raw = orc.head_raw(orc.pafpn({"dark3": f[0], "dark4": f[1], "dark5": f[2]}, s2, True), s2, True)
    tr = orc.TRACE; orc.TRACE = None
    orc.EMULATE_STORAGE = cno.EMULATE_STORAGE = False
    return tr, f
ref, fr = trace(False); emu, fe = trace(True)
for k, name in zip(("dark3", "dark4", "dark5"), (0, 1, 2)):
    t = eng.yx.features[k].tensor().float().permute(0, 3, 1, 2).cpu()
    print("feature %s: engine err %.5f emu err %.5f (mean |ref| %.4f)" % (k, (t - fr[name]).abs().mean(), (fe[name] - fr[name]).abs().mean(), fr[name].abs().mean()))
lo, hi = eng.range
for op in eng.yx.ops[lo:hi]:
    if not isinstance(op, ConvOp):
        continue
    for hd in op.heads:
        if hd.prefix not in ref:
            continue
        t = hd.out.tensor().float().permute(0, 3, 1, 2).cpu()
        r, e = ref[hd.prefix], emu[hd.prefix]
        print("%-28s %s engine err %.5f | emu err %.5f | mean |ref| %.4f" % (hd.prefix, tuple(r.shape[1:]), (t - r).abs().mean(), (e - r).abs().mean(), r.abs().mean()))
