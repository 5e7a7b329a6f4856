"""diagnostic: YoloxEngine at width 0.75 (channel counts 48/96/192/384/768: non-power-of-two vectors) against the oracle, worst gradients"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import yolox_oracle as orc
from yolov7_d2_b200.engine import YoloxEngine

dev = torch.device("cuda:0")
width = float(sys.argv[1]) if len(sys.argv) > 1 else 0.75
torch.set_num_threads(16)
sd = orc.yolox_state_dict(3, width=width)
images, labels = orc.synthetic_batch(4, 256, 5, max_gt=6)
eng = YoloxEngine(4, 256, 256, width_mul=width, device=dev)
eng.load_state_dict(sd)
eng.images_u8.copy_(images.to(dev)); eng.labels.copy_(labels.to(dev))
eng.pack_weights(); eng.preprocess(); eng.forward_features(True)
gen = torch.Generator().manual_seed(77)
n, a, ch = eng.outputs.shape
g_raw = (torch.randn(n, a, ch, generator=gen) * 1e-2).to(torch.bfloat16).float()
for k, (h, w, s, a_off) in enumerate(eng.levels):
    gl = g_raw[:, a_off:a_off + h * w]
    eng.d_cls[k].copy_(gl[..., 5:].reshape(n, h, w, ch - 5).to(dev))
    eng.d_ro[k].zero_(); eng.d_ro[k][..., :5].copy_(gl[..., :5].reshape(n, h, w, 5).to(dev))
    eng.bias_acc[k].copy_(gl.double().sum((0, 1)).to(dev))
eng.backward(); torch.cuda.synchronize()
def oracle(emulate):
    orc.EMULATE_STORAGE = emulate
    s2 = {k: v.clone() for k, v in sd.items()}
    for k, v in s2.items():
        if v.dtype == torch.float32 and "running" not in k: v.requires_grad_(True)
    raw = orc.head_raw(orc.pafpn(orc.csp_darknet(images.float(), s2, True), s2, True), s2, True)
    flat = torch.cat([r.permute(0, 2, 3, 1).reshape(r.shape[0], -1, r.shape[1]) for r in raw], 1)
    (flat * g_raw).sum().backward()
    orc.EMULATE_STORAGE = False
    return {k: v.grad for k, v in s2.items() if v.requires_grad}
ref, emu = oracle(False), oracle(True)
rows = []
for name in eng.param_names:
    g = eng.grads[name].cpu().flatten().double(); r = ref[name].flatten().double(); e = emu[name].flatten().double()
    rows.append((float((g @ r) / (g.norm() * r.norm() + 1e-30)), float((e @ r) / (e.norm() * r.norm() + 1e-30)), float(g.norm() / (r.norm() + 1e-30)), name))
rows.sort()
for r in rows[:15]: print("  %.4f (emu %.4f) ratio %.3f %s" % r)
