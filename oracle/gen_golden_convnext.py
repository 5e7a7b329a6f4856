"""TEST INFRASTRUCTURE -- generates tests/golden/convnext.npz from the UNMODIFIED reference ConvNeXt classes.

Run in the build container (where /root/reference exists):   python -m oracle.gen_golden_convnext
The reference file is imported through oracle/ref_shim.py (timm's `trunc_normal_` / `DropPath` and alfred's logger are
stubbed: they only matter for initialisation and for stochastic depth, which parity runs switch off -- SURVEY.md par.8a C1).
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

from . import ref_shim
from . import convnext_oracle as cnx

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "convnext.npz")


def load_reference():
    ref_shim.install()

    class DropPath(torch.nn.Module):  # timm.models.layers.DropPath
        def __init__(self, p=0.0):
            super().__init__()
            self.p = p

        def forward(self, x):
            if self.p == 0.0 or not self.training:
                return x
            keep = 1 - self.p
            mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            return x * mask / keep

    tm = types.ModuleType("timm")
    tm.__path__ = []
    sys.modules["timm"] = tm
    tmm = types.ModuleType("timm.models")
    tmm.__path__ = []
    sys.modules["timm.models"] = tmm
    lay = types.ModuleType("timm.models.layers")
    lay.trunc_normal_ = torch.nn.init.trunc_normal_
    lay.DropPath = DropPath
    sys.modules["timm.models.layers"] = lay
    return importlib.import_module("yolov7.modeling.backbone.convnext")


def _np(t):
    return t.detach().cpu().numpy()


def main():
    mod = load_reference()
    out = {}
    torch.manual_seed(0)

    # ---- one Block, dim 32, trained-like parameters: forward + all gradients ----
    dim = 32
    blk = mod.Block(dim=dim, drop_path=0.0, layer_scale_init_value=1e-6)
    sd = cnx.convnext_state_dict(3, depths=(1, 0, 0, 0), dims=(dim, 8, 8, 8), trained_like=True)
    bsd = {k[len("stages.0.0."):]: v for k, v in sd.items() if k.startswith("stages.0.0.")}
    blk.load_state_dict(bsd, strict=True)
    x = torch.randn(2, dim, 12, 20, requires_grad=True)
    y = blk(x)
    gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(5))
    y.backward(gy)
    out["block_x"], out["block_y"], out["block_gy"], out["block_gx"] = _np(x), _np(y), _np(gy), _np(x.grad)
    for k, v in bsd.items():
        out["block_sd/" + k] = _np(v)
    for k, p in blk.named_parameters():
        out["block_grad/" + k] = _np(p.grad)

    # ---- LayerNorm channels_first ----
    ln = mod.LayerNorm(24, eps=1e-6, data_format="channels_first")
    with torch.no_grad():
        ln.weight.copy_(torch.rand(24) + 0.5)
        ln.bias.copy_(torch.rand(24) - 0.5)
    xl = torch.randn(2, 24, 5, 7)
    out["ln_x"], out["ln_w"], out["ln_b"], out["ln_y"] = _np(xl), _np(ln.weight), _np(ln.bias), _np(ln(xl))

    # ---- tiny ConvNeXt (depths 1,1,2,1; dims 16,32,48,64), trained-like parameters: forward_features + input/param gradients ----
    depths, dims = (1, 1, 2, 1), (16, 32, 48, 64)
    net = mod.ConvNeXt(in_chans=3, depths=list(depths), dims=list(dims), drop_path_rate=0.0, layer_scale_init_value=1e-6, out_indices=[0, 1, 2, 3])
    nsd = cnx.convnext_state_dict(7, depths=depths, dims=dims, trained_like=True)
    net.load_state_dict(nsd, strict=True)
    img = cnx.synthetic_images(2, 64, seed=11).float()
    img.requires_grad_(True)
    feats = net(img)
    gens = torch.Generator().manual_seed(13)
    gouts = [torch.randn(f.shape, generator=gens) for f in feats]
    sum((f * g).sum() for f, g in zip(feats, gouts)).backward()
    out["net_img"] = _np(img).astype(np.uint8)
    for i, (f, g) in enumerate(zip(feats, gouts)):
        out[f"net_out{i}"], out[f"net_gout{i}"] = _np(f), _np(g)
    out["net_gimg"] = _np(img.grad)
    for k, p in net.named_parameters():
        out["net_grad/" + k] = _np(p.grad)
    out["net_depths"], out["net_dims"] = np.array(depths), np.array(dims)

    # ---- default initialisation statistics of the reference (ConvNeXt._init_weights): names + shapes of ConvNeXt-T ----
    full = mod.ConvNeXt(in_chans=3, depths=[3, 3, 9, 3], dims=[96, 192, 384, 768], drop_path_rate=0.0, layer_scale_init_value=1e-6, out_indices=[0, 1, 2, 3])
    names = sorted(full.state_dict().keys())
    out["tiny_names"] = np.array(names)
    out["tiny_shapes"] = np.array([",".join(str(d) for d in full.state_dict()[n].shape) for n in names])
    out["tiny_params"] = np.array(sum(p.numel() for p in full.parameters()))

    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "%.1f KB" % (os.path.getsize(OUT) / 1e3))


if __name__ == "__main__":
    main()
