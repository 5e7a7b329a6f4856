"""TEST INFRASTRUCTURE -- generates tests/golden/sparseinst.npz from the UNMODIFIED reference BaseIAMDecoder
(yolov7/modeling/transcoders/decoder_sparseinst.py; fvcore.nn.weight_init and detectron2's Registry / Conv2d are stubbed).
Run in the build container:   python -m oracle.gen_golden_sparseinst"""
import importlib
import os
import sys
import types

import numpy as np
import torch

from . import ref_shim
from . import sparseinst_oracle as sio

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sparseinst.npz")


def ns(**kw):
    return types.SimpleNamespace(**kw)


def load_reference():
    ref_shim.install()
    wi = types.ModuleType("fvcore.nn.weight_init")
    wi.c2_msra_fill = lambda m: torch.nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
    wi.c2_xavier_fill = lambda m: torch.nn.init.kaiming_uniform_(m.weight, a=1)
    for name in ("fvcore", "fvcore.nn"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    sys.modules["fvcore.nn.weight_init"] = wi

    class Registry(dict):
        def __init__(self, name):
            super().__init__()

        def register(self, obj=None):
            if obj is None:
                return lambda o: self.register(o)
            self[obj.__name__] = obj
            return obj

    reg = types.ModuleType("detectron2.utils.registry")
    reg.Registry = Registry
    u = sys.modules.get("detectron2.utils") or types.ModuleType("detectron2.utils")
    u.__path__ = []
    sys.modules["detectron2.utils"], sys.modules["detectron2.utils.registry"] = u, reg
    sys.modules["detectron2.layers"].Conv2d = torch.nn.Conv2d
    pk = types.ModuleType("yolov7.modeling.transcoders")
    pk.__path__ = [os.path.join(ref_shim.REF, "yolov7", "modeling", "transcoders")]
    sys.modules["yolov7.modeling.transcoders"] = pk
    return importlib.import_module("yolov7.modeling.transcoders.decoder_sparseinst")


def main():
    mod = load_reference()
    dim, nm, kd, nc, convs, cin = 64, 20, 32, 8, 2, 30
    cfg = ns(MODEL=ns(SPARSE_INST=ns(ENCODER=ns(NUM_CHANNELS=cin), DECODER=ns(SCALE_FACTOR=2.0, OUTPUT_IAM=False, NUM_MASKS=nm, KERNEL_DIM=kd, NUM_CLASSES=nc,
                                                                                  INST=ns(DIM=dim, CONVS=convs), MASK=ns(DIM=dim, CONVS=convs)))))
    dec = mod.BaseIAMDecoder(cfg)
    sd = sio.decoder_state_dict(5, in_channels=cin, dim=dim, num_masks=nm, kernel_dim=kd, num_classes=nc, num_convs=convs)
    dec.load_state_dict(sd, strict=True)
    dec.eval()
    g = torch.Generator().manual_seed(6)
    feat = torch.randn(2, cin, 12, 20, generator=g)
    with torch.no_grad():
        out = dec(feat)
        x = torch.cat([dec.compute_coordinates(feat), feat], 1)
        logits, kernel, scores, iam = dec.inst_branch(x)
    res = {"feat": feat.numpy(), "pred_logits": out["pred_logits"].numpy(), "pred_masks": out["pred_masks"].numpy(), "pred_scores": out["pred_scores"].numpy(),
           "pred_kernel": kernel.numpy(), "iam": iam.numpy(), "dims": np.array([dim, nm, kd, nc, convs, cin])}
    np.savez_compressed(OUT, **res)
    print("wrote", OUT, "%.1f KB" % (os.path.getsize(OUT) / 1e3))
    # GroupIAMDecoder (decoder_sparseinst.py:172-250): 4 groups
    groups = 4
    cfg.MODEL.SPARSE_INST.DECODER.GROUPS = groups
    gdec = mod.GroupIAMDecoder(cfg)
    gsd = sio.decoder_state_dict(7, in_channels=cin, dim=dim, num_masks=nm, kernel_dim=kd, num_classes=nc, num_convs=convs, groups=groups)
    gdec.load_state_dict(gsd, strict=True)
    gdec.eval()
    with torch.no_grad():
        gout = gdec(feat)
        x = torch.cat([gdec.compute_coordinates(feat), feat], 1)
        _, gkernel, _, giam = gdec.inst_branch(x)
    gres = {"feat": feat.numpy(), "pred_logits": gout["pred_logits"].numpy(), "pred_masks": gout["pred_masks"].numpy(), "pred_scores": gout["pred_scores"].numpy(),
            "pred_kernel": gkernel.numpy(), "iam": giam.numpy(), "dims": np.array([dim, nm, kd, nc, convs, cin, groups])}
    gpath = OUT.replace("sparseinst.npz", "sparseinst_group.npz")
    np.savez_compressed(gpath, **gres)
    print("wrote", gpath, "%.1f KB" % (os.path.getsize(gpath) / 1e3))


if __name__ == "__main__":
    main()
