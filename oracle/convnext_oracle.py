"""TEST INFRASTRUCTURE -- CPU restatement (plain torch fp32) of the reference's ConvNeXt backbone path.

Pinned by tests/golden/convnext.npz, which oracle/gen_golden_convnext.py produced from the *unmodified* reference classes
(yolov7/modeling/backbone/convnext.py imported through oracle/ref_shim.py); tests/test_oracle_golden.py re-checks this file
against those vectors on every CPU run.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it.

Every function is functional (explicit `sd` = reference-layout state_dict) and cites the reference lines it restates.
`EMULATE_STORAGE` rounds the tensors that the CUDA path stores in 16 bit (same yardstick idea as oracle/yolox_oracle.py).
"""
import math

import torch
import torch.nn.functional as F

EMULATE_STORAGE = False
LN_EPS = 1e-6  # convnext.py:43,85,90,109


def _q(t):
    return t.to(torch.bfloat16).to(torch.float32) if EMULATE_STORAGE else t


def layer_norm_channels(x_nchw, weight, bias, eps=LN_EPS):
    """LayerNorm(data_format="channels_first"), convnext.py:201-206: per pixel over C, biased variance."""
    u = x_nchw.mean(1, keepdim=True)
    s = (x_nchw - u).pow(2).mean(1, keepdim=True)
    x = (x_nchw - u) / torch.sqrt(s + eps)
    return weight[:, None, None] * x + bias[:, None, None]


def block_forward(x, sd, prefix, drop_path=None):
    """Block.forward, convnext.py:47-60 (drop_path: optional per-sample keep mask / keep_prob tensor [N], timm DropPath)."""
    c = x.shape[1]
    inp = x
    x = F.conv2d(x, sd[prefix + "dwconv.weight"], sd[prefix + "dwconv.bias"], padding=3, groups=c)  # :49
    x = x.permute(0, 2, 3, 1)                                                                              # :50
    x = _q(F.layer_norm(x, (c,), sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], LN_EPS))            # :51, :198-199
    x = F.linear(x, _q(sd[prefix + "pwconv1.weight"]), sd[prefix + "pwconv1.bias"])                        # :52
    x = _q(F.gelu(_q(x)))                                                                                   # :53 (exact erf GELU)
    x = F.linear(x, _q(sd[prefix + "pwconv2.weight"]), sd[prefix + "pwconv2.bias"])                        # :54
    g = sd.get(prefix + "gamma")
    if g is not None:
        x = g * x                                                                                           # :55-56
    x = x.permute(0, 3, 1, 2)                                                                              # :57
    if drop_path is not None:
        x = x * drop_path.view(-1, 1, 1, 1)
    return _q(inp + x)                                                                                      # :59


def forward_features(x, sd, depths=(3, 3, 9, 3), out_indices=(0, 1, 2, 3), prefix=""):
    """ConvNeXt.forward_features, convnext.py:149-159 with the layers built at :80-117.  x: [N,3,H,W] fp32."""
    outs = []
    for i in range(4):
        p = f"{prefix}downsample_layers.{i}."
        if i == 0:   # stem: Conv 4x4 s4 then LN(channels_first)   :81-85
            x = F.conv2d(_q(x), _q(sd[p + "0.weight"]), sd[p + "0.bias"], stride=4)
            x = _q(layer_norm_channels(x, sd[p + "1.weight"], sd[p + "1.bias"]))
        else:        # LN(channels_first) then Conv 2x2 s2              :86-91
            x = _q(layer_norm_channels(x, sd[p + "0.weight"], sd[p + "0.bias"]))
            x = _q(F.conv2d(x, _q(sd[p + "1.weight"]), sd[p + "1.bias"], stride=2))
        for j in range(depths[i]):
            x = block_forward(x, sd, f"{prefix}stages.{i}.{j}.")
        if i in out_indices:
            outs.append(layer_norm_channels(x, sd[f"{prefix}norm{i}.weight"], sd[f"{prefix}norm{i}.bias"]))
    return tuple(outs)


def convnext_state_dict(seed=0, in_chans=3, depths=(3, 3, 9, 3), dims=(96, 192, 384, 768), layer_scale=1e-6, trained_like=False):
    """Reference-layout state_dict (names of convnext.py:76-117).  Default values follow ConvNeXt._init_weights (:119-122:
    trunc_normal(std .02) weights, zero biases, LN weight 1 / bias 0, gamma = layer_scale).  trained_like=True draws
    biases, LN affine parameters and gamma from wider distributions so that every term of the block matters in a parity test."""
    g = torch.Generator().manual_seed(seed)

    def tn(*shape, std=0.02):
        t = torch.empty(*shape)
        torch.nn.init.trunc_normal_(t, std=std, generator=g)
        return t

    def rnd(*shape, lo, hi):
        return torch.rand(*shape, generator=g) * (hi - lo) + lo

    sd = {}

    def ln(name, c):
        sd[name + ".weight"] = rnd(c, lo=0.5, hi=1.5) if trained_like else torch.ones(c)
        sd[name + ".bias"] = rnd(c, lo=-0.3, hi=0.3) if trained_like else torch.zeros(c)

    def conv(name, co, ci, k, std=0.02):
        sd[name + ".weight"] = tn(co, ci, k, k, std=std)
        sd[name + ".bias"] = rnd(co, lo=-0.2, hi=0.2) if trained_like else torch.zeros(co)

    wstd = 0.02
    conv("downsample_layers.0.0", dims[0], in_chans, 4, std=0.1 if trained_like else wstd)
    ln("downsample_layers.0.1", dims[0])
    for i in range(3):
        ln(f"downsample_layers.{i + 1}.0", dims[i])
        conv(f"downsample_layers.{i + 1}.1", dims[i + 1], dims[i], 2, std=(1.0 / math.sqrt(4 * dims[i])) if trained_like else wstd)
    for i in range(4):
        c = dims[i]
        for j in range(depths[i]):
            p = f"stages.{i}.{j}."
            sd[p + "dwconv.weight"] = tn(c, 1, 7, 7, std=0.15 if trained_like else wstd)
            sd[p + "dwconv.bias"] = rnd(c, lo=-0.2, hi=0.2) if trained_like else torch.zeros(c)
            ln(p + "norm", c)
            sd[p + "pwconv1.weight"] = tn(4 * c, c, std=(1.0 / math.sqrt(c)) if trained_like else wstd)
            sd[p + "pwconv1.bias"] = rnd(4 * c, lo=-0.5, hi=0.5) if trained_like else torch.zeros(4 * c)
            sd[p + "pwconv2.weight"] = tn(c, 4 * c, std=(1.0 / math.sqrt(4 * c)) if trained_like else wstd)
            sd[p + "pwconv2.bias"] = rnd(c, lo=-0.2, hi=0.2) if trained_like else torch.zeros(c)
            if layer_scale > 0:
                sd[p + "gamma"] = rnd(c, lo=0.05, hi=0.6) if trained_like else layer_scale * torch.ones(c)
    for i in range(4):
        ln(f"norm{i}", dims[i])
    return sd


from yolov7_d2_b200.synth import synthetic_images  # noqa: E402,F401
