"""TEST INFRASTRUCTURE -- makes the reference's hot-path Python files importable from /root/reference.

Only `oracle/gen_golden.py` (run in the build container, where /root/reference exists) uses this module; nothing
in the product, the GPU tests, smoke() or bench.py may import it: /root/reference does not exist on the GPU box.

The reference needs detectron2, pycocotools, omegaconf, alfred ... none of which are installed here.  We register
minimal stand-ins in `sys.modules` (a dict registry, ShapeSpec, Backbone) and declare the `yolov7.*` packages as
namespace stubs so that their `__init__.py` files (which import every architecture of the zoo) are not executed;
the individual hot-path files are then imported *unmodified* from the reference tree.
"""
import importlib
import os
import sys
import types

import torch.nn as nn

REF = os.environ.get("YB200_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "yolov7", "modeling"))


class _Registry(dict):
    def register(self, obj=None):
        if obj is None:
            return lambda o: self.register(o)
        self[obj.__name__] = obj
        return obj

    def get(self, name):
        return self[name]


class _ShapeSpec:
    def __init__(self, channels=None, height=None, width=None, stride=None):
        self.channels, self.height, self.width, self.stride = channels, height, width, stride


class _Backbone(nn.Module):
    @property
    def size_divisibility(self):
        return 0


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    backbone_reg, arch_reg = _Registry(), _Registry()

    def get_norm(norm, ch):
        return nn.BatchNorm2d(ch) if norm == "BN" else None

    _mod("detectron2").__path__ = []
    _mod("detectron2.layers", ShapeSpec=_ShapeSpec, get_norm=get_norm).__path__ = []
    _mod("detectron2.layers.batch_norm", get_norm=get_norm)
    _mod("detectron2.modeling", META_ARCH_REGISTRY=arch_reg).__path__ = []
    _mod("detectron2.modeling.backbone", Backbone=_Backbone, BACKBONE_REGISTRY=backbone_reg).__path__ = []
    _mod("detectron2.modeling.backbone.build", BACKBONE_REGISTRY=backbone_reg)
    _mod("omegaconf", base=None)
    pc = _mod("pycocotools")
    pc.__path__ = []
    pc.mask = _mod("pycocotools.mask")
    _mod("alfred", logger=types.SimpleNamespace(info=lambda *a, **k: None, warning=lambda *a, **k: None))
    for sub in ("", ".utils", ".modeling", ".modeling.backbone", ".modeling.backbone.layers", ".modeling.head", ".modeling.neck"):
        _pkg("yolov7" + sub, os.path.join(REF, "yolov7", *[p for p in sub.split(".") if p]))
    _installed = True


def load():
    """returns (boxes, darknetx, yolo_pafpn, yolox_head, wrappers, checkpoint) reference modules"""
    install()
    names = (
        "yolov7.utils.boxes",
        "yolov7.modeling.backbone.darknetx",
        "yolov7.modeling.neck.yolo_pafpn",
        "yolov7.modeling.head.yolox_head",
        "yolov7.modeling.backbone.layers.wrappers",
        "yolov7.utils.checkpoint",
    )
    return tuple(importlib.import_module(n) for n in names)
