"""SparseInst IAM decoder forward at the shipped size (BASELINE.json configs[4] per-GPU share: 16 images, 80x80 map, 256+2 channels, 100 masks)."""
import json, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolov7_d2_b200.sparseinst import BaseIAMDecoder

ns = lambda **kw: types.SimpleNamespace(**kw)
cfg = ns(MODEL=ns(SPARSE_INST=ns(ENCODER=ns(NUM_CHANNELS=256), DECODER=ns(SCALE_FACTOR=2.0, OUTPUT_IAM=False, NUM_MASKS=100, KERNEL_DIM=128, NUM_CLASSES=80,
                                                                              INST=ns(DIM=256, CONVS=4), MASK=ns(DIM=256, CONVS=4)))))
dec = BaseIAMDecoder(cfg)
feat = torch.randn(16, 256, 80, 80, device="cuda")
for _ in range(3):
    dec(feat)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    dec(feat)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
gf = 16 * (8 * 7.55 + 2 * 6400 * 9 * 256 * 100 / 1e9 + 0.33 + 0.16 + 2 * 6400 * 256 * 128 / 1e9)  # 8 3x3 convs + iam conv + bmm's + projection
print(json.dumps({"workload": "SparseInst BaseIAMDecoder forward, 16 x 258 x 80 x 80, 100 masks (eager launches)", "ms": ms, "images_per_s": 16 / ms * 1e3, "tflops": gf / ms}))
