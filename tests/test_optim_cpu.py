"""Host logic of the flat optimizer (no GPU): the segment table reproduces the reference's parameter groups
(yolov7/optimizer/build.py:77-170) and the optimizer refuses to run without CUDA buffers."""
import pytest
import torch

from yolov7_d2_b200 import capi, optim


def _reference_groups(named, weight_decay, wd_norm, wd_bias, bias_lr_factor, overrides):
    """what get_optimizer_param_groups_lr + _weight_decay assign to each parameter (restated from build.py:77-170)"""
    out = {}
    for name, is_norm in named:
        module_name, _, pname = name.rpartition(".")
        lr = 1.0 * (bias_lr_factor if pname == "bias" else 1.0)
        for k, m in overrides.items():
            if k in module_name:
                lr *= m
        wd = weight_decay
        if is_norm:
            wd = weight_decay if wd_norm is None else wd_norm
        elif pname == "bias":
            wd = weight_decay if wd_bias is None else wd_bias
        out[name] = (wd, lr)
    return out


def test_segments_match_reference_groups():
    layout = [("backbone.stem.conv.weight", 0, 30), ("backbone.stem.bn.weight", 32, 4), ("backbone.stem.bn.bias", 36, 4),
              ("head.cls_preds.0.weight", 40, 17), ("head.cls_preds.0.bias", 60, 3)]
    segs = optim.param_segments(layout, 64, 5e-4, 0.0, 1e-5, 2.0, {"backbone": 0.1})
    ref = _reference_groups([(n, ".bn." in n) for n, _, _ in layout], 5e-4, 0.0, 1e-5, 2.0, {"backbone": 0.1})
    begins = [s[0] for s in segs]
    assert begins == sorted(begins) and begins[0] == 0

    def lookup(i):
        s = max(j for j, b in enumerate(begins) if b <= i)
        return segs[s][1], segs[s][2]

    for name, off, n in layout:
        for i in (off, off + n - 1):
            wd, lr = lookup(i)
            assert wd == pytest.approx(ref[name][0]) and lr == pytest.approx(ref[name][1]), name
    # alignment gaps are frozen
    for i in (30, 31, 57, 59, 63):
        assert lookup(i) == (0.0, 0.0)


def test_adjacent_equal_segments_merge():
    layout = [("a.conv.weight", 0, 8), ("b.conv.weight", 8, 8), ("a.bn.weight", 16, 4), ("a.bn.bias", 20, 4)]
    segs = optim.param_segments(layout, 24, 1e-4, 0.0)
    assert segs == [(0, 1e-4, 1.0), (16, 0.0, 1.0)]


def test_no_cpu_implementation():
    p = torch.zeros(8)
    with pytest.raises(capi.Yb200Error):
        optim.FlatOptimizer(p, torch.zeros(8), [(0, 0.0, 1.0)], 0.1, "sgd")


def test_unknown_optimizer_name():
    class S:
        OPTIMIZER = "lamb"

    class C:
        SOLVER = S()

    with pytest.raises(KeyError):
        optim.build_optimizer_mapper(C(), object())
