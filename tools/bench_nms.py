"""NMS throughput on the clustered stress set ([64, 8400, 85], conf 0.001, IoU 0.65) -- the `nms` key of bench.py, alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolov7_d2_b200 import synth
from yolov7_d2_b200.modeling import postprocess

dev = torch.device("cuda:0")
pred = synth.clustered_predictions(4, 8400, 80, 7).repeat(16, 1, 1).to(dev)
work = [pred.clone() for _ in range(12)]
for w in work[:2]:
    postprocess(w, 80, 0.001, 0.65)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for w in work[2:]:
    dets = postprocess(w, 80, 0.001, 0.65)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("postprocess: %.3f ms per [64, 8400, 85] call = %.1f M boxes/s; detections kept in image 0: %d" % (ms, 64 * 8400 / ms / 1e3, 0 if dets[0] is None else dets[0].shape[0]))
