"""Host mirror of the reference's optimizer builders (yolov7/optimizer/build.py) on top of the flat fp32 buffers.

`sgd(cfg, model)` / `adamw(cfg, model)` / `build_optimizer_mapper(cfg, model)` keep the reference's names and the meaning of
the `cfg.SOLVER.*` keys they read (build.py:24-60, 234-256, 294-300).  The returned object is a `torch.optim.Optimizer`
(LR schedulers and detectron2's trainer only touch `param_groups[i]["lr"]`, `step()` and `zero_grad()`), but `step()` is ONE
kernel launch over the whole model (csrc/optim.cu) instead of a Python loop over ~250 tensors: per-parameter weight decay
and learning-rate multipliers are a segment table in device memory.  There is no CPU implementation.
"""
import ctypes

import torch

from . import capi

# Parameters of torch normalisation modules (build.py:96-110 tests isinstance against BatchNorm / LayerNorm / GroupNorm ...).  In the
# YOLOX plan those are exactly the `.bn.` tensors.  ConvNeXt's LayerNorm is a custom nn.Module (convnext.py:182-206), NOT in the
# reference's norm_module_types: its weight takes WEIGHT_DECAY and its bias WEIGHT_DECAY_BIAS, like any other parameter.  An engine
# can override the rule by exposing `norm_param_names` (a set of parameter names).
_NORM_SUFFIXES = (".bn.weight", ".bn.bias")


def _solver(cfg, key, default):
    solver = getattr(cfg, "SOLVER", None)
    if solver is None:
        return default
    if isinstance(solver, dict):
        return solver.get(key, default)
    return getattr(solver, key, default)


def param_segments(param_layout, total, weight_decay, weight_decay_norm=None, weight_decay_bias=None, bias_lr_factor=1.0,
                   lr_multipliers_overwrite=None, norm_param_names=None):
    """[(begin, weight_decay, lr_multiplier)] covering [0, total) for parameters placed at `param_layout` = [(name, offset, numel)]
    (ascending offsets; gaps are alignment padding and get lr multiplier 0, i.e. they are never modified).

    Follows get_optimizer_param_groups_lr / _weight_decay (build.py:77-170): normalisation-layer parameters take
    weight_decay_norm, parameters *named* "bias" take weight_decay_bias and lr * bias_lr_factor, and a multiplier applies
    to every parameter whose module name contains one of the override keys.  Adjacent equal segments are merged
    (the reference's reduce_param_groups)."""
    wd_norm = weight_decay if weight_decay_norm is None else weight_decay_norm
    wd_bias = weight_decay if weight_decay_bias is None else weight_decay_bias
    segs = []

    def push(off, wd, mult):
        if not segs or (segs[-1][1], segs[-1][2]) != (wd, mult):
            segs.append((off, float(wd), float(mult)))

    end = 0
    for name, off, n in param_layout:
        assert off >= end, "param_layout must be sorted and non-overlapping"
        if off > end:
            push(end, 0.0, 0.0)
        module_name, _, pname = name.rpartition(".")
        wd, mult = weight_decay, 1.0
        if (name in norm_param_names) if norm_param_names is not None else name.endswith(_NORM_SUFFIXES):
            wd = wd_norm
        elif pname == "bias":
            wd = wd_bias
        if pname == "bias":
            mult *= bias_lr_factor
        for key, m in (lr_multipliers_overwrite or {}).items():
            if key in module_name:
                mult *= m
        push(off, wd, mult)
        end = off + n
    if end < total:
        push(end, 0.0, 0.0)
    if not segs:
        segs.append((0, 0.0, 0.0))
    return segs


class FlatOptimizer(torch.optim.Optimizer):
    """one param group holding the flat parameter tensor; `param_groups[0]["lr"]` is what schedulers drive"""

    def __init__(self, flat_param, flat_grad, segments, lr, kind, momentum=0.0, dampening=0.0, nesterov=False, betas=(0.9, 0.999), eps=1e-8,
                 clip_norm=0.0, grad_scale=1.0):
        if not flat_param.is_cuda:
            raise capi.Yb200Error("FlatOptimizer needs the engine's CUDA buffers; the hot path has no CPU implementation")
        assert flat_param.dtype == torch.float32 and flat_param.is_contiguous() and flat_grad.shape == flat_param.shape
        self.flat_param, self.flat_grad = flat_param, flat_grad
        self.kind = kind
        dev = flat_param.device
        self.seg_begin = torch.tensor([s[0] for s in segments], dtype=torch.int64, device=dev)
        self.seg_wd = torch.tensor([s[1] for s in segments], dtype=torch.float32, device=dev)
        self.seg_lr = torch.tensor([s[2] for s in segments], dtype=torch.float32, device=dev)
        self.nseg = len(segments)
        self.clip_norm = float(clip_norm)
        self.grad_scale = float(grad_scale)
        L = capi.lib()
        self.norm_ws = torch.empty(int(L.yb200_grad_norm_workspace()), dtype=torch.uint8, device=dev)
        self.total_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, nesterov=nesterov, betas=betas, eps=eps)
        super().__init__([flat_param], defaults)
        # The optimizer state lives in torch's own `self.state` (names as torch.optim.SGD / AdamW), so state_dict() /
        # load_state_dict() -- detectron2's checkpointer -- carry momentum / moments / the step count across a resume.
        st = self.state[flat_param]
        st["step"] = 0
        if kind == "adamw":
            st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(flat_param), torch.zeros_like(flat_param)
        elif momentum != 0.0:
            st["momentum_buffer"] = torch.zeros_like(flat_param)
        # torch.cuda.amp.GradScaler finds, unscales and inf-checks gradients through `param.grad`: the flat parameter's gradient IS
        # the flat gradient buffer, so scaler.unscale_() / scaler.step() / scaler.update() work unchanged (SOLVER.AMP.ENABLED configs).
        flat_param.grad = flat_grad

    @property
    def steps(self):
        return int(self.state[self.flat_param]["step"])

    @property
    def state_a(self):
        st = self.state[self.flat_param]
        return st.get("exp_avg") if self.kind == "adamw" else st.get("momentum_buffer")

    @property
    def state_b(self):
        return self.state[self.flat_param].get("exp_avg_sq")

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        L, sp = capi.lib(), capi.stream_ptr()
        g = self.param_groups[0]
        n = self.flat_param.numel()
        f = ctypes.c_float
        norm = None
        if self.clip_norm > 0:
            capi.check(L.yb200_grad_norm(capi.ptr(self.flat_grad), ctypes.c_int64(n), f(self.grad_scale), capi.ptr(self.norm_ws), capi.ptr(self.total_norm),
                                         sp), "grad_norm")
            norm = self.total_norm
        if self.flat_param.grad is None or self.flat_param.grad.data_ptr() != self.flat_grad.data_ptr():
            self.flat_param.grad = self.flat_grad  # zero_grad(set_to_none=True) / load_state_dict dropped the alias
        st = self.state[self.flat_param]
        st["step"] = int(st["step"]) + 1
        if self.kind == "sgd":
            capi.check(L.yb200_sgd_step(capi.ptr(self.flat_param), capi.ptr(self.flat_grad), capi.ptr(self.state_a), ctypes.c_int64(n),
                                        capi.ptr(self.seg_begin), capi.ptr(self.seg_wd), capi.ptr(self.seg_lr), self.nseg, f(g["lr"]),
                                        f(g["momentum"]), f(g["dampening"]), int(bool(g["nesterov"])), int(self.steps == 1), f(self.grad_scale),
                                        capi.ptr(norm), f(self.clip_norm), sp), "sgd_step")
        else:
            capi.check(L.yb200_adamw_step(capi.ptr(self.flat_param), capi.ptr(self.flat_grad), capi.ptr(self.state_a), capi.ptr(self.state_b),
                                          ctypes.c_int64(n), capi.ptr(self.seg_begin), capi.ptr(self.seg_wd), capi.ptr(self.seg_lr), self.nseg,
                                          f(g["lr"]), f(g["betas"][0]), f(g["betas"][1]), f(g["eps"]), self.steps, f(self.grad_scale), capi.ptr(norm),
                                          f(self.clip_norm), sp), "adamw_step")
        return loss

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        st = self.state[self.flat_param]
        for k in ("momentum_buffer", "exp_avg", "exp_avg_sq"):
            if k in st and (st[k].shape != self.flat_param.shape or st[k].dtype != torch.float32 or not st[k].is_contiguous()):
                raise capi.Yb200Error(f"optimizer state {k} does not match the flat parameter buffer")
        st["step"] = int(st.get("step", 0))


class FlatOptimizerGroup(torch.optim.Optimizer):
    """Several flat buffers (e.g. ConvNeXt backbone + YOLOX neck / head plans) behind ONE torch.optim.Optimizer: param group i is buffer i, so LR
    schedulers drive every buffer through `param_groups[i]["lr"]`; step() is one fused launch per buffer."""

    def __init__(self, children):
        self.children = list(children)
        d = dict(self.children[0].defaults)
        super().__init__([{"params": [c.flat_param]} for c in self.children], d)
        for c, g in zip(self.children, self.param_groups):
            g.update({k: v for k, v in c.param_groups[0].items() if k != "params"})

    @property
    def grad_scale(self):
        return self.children[0].grad_scale

    @grad_scale.setter
    def grad_scale(self, v):
        for c in self.children:
            c.grad_scale = float(v)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for c, g in zip(self.children, self.param_groups):
            c.param_groups[0].update({k: v for k, v in g.items() if k != "params"})
            c.step()
        return loss

    def zero_grad(self, set_to_none=False):
        for c in self.children:
            c.zero_grad()

    def state_dict(self):
        return {"children": [c.state_dict() for c in self.children]}

    def load_state_dict(self, sd):
        for c, d in zip(self.children, sd["children"]):
            c.load_state_dict(d)
        for c, g in zip(self.children, self.param_groups):
            g.update({k: v for k, v in c.param_groups[0].items() if k != "params"})


def _flat_buffers(model):
    eng = getattr(model, "engine", None) or getattr(getattr(model, "module", None), "engine", None) or model
    if not hasattr(eng, "flat_param") and not hasattr(eng, "flat_buffers"):
        raise capi.Yb200Error("optimizer needs a model that exposes the engine's flat_param / flat_grad / param_layout")
    if type(model).__name__ == "DistributedDataParallel":
        raise capi.Yb200Error("the flat optimizer writes gradients straight into the flat buffer, so torch DDP's per-parameter reducer hooks never "
                              "fire: do not wrap the model in DistributedDataParallel; all-reduce `model.engine.flat_grad` instead "
                              "(yolov7_d2_b200.dist.GradientBuckets, INTEGRATION.md par.4)")
    inner = getattr(model, "module", model)
    if hasattr(inner, "attach_flat_grads"):
        inner.attach_flat_grads()
    return eng


def _clip_norm(cfg):
    clip = _solver(cfg, "CLIP_GRADIENTS", None)
    if clip is None:
        return 0.0
    get = clip.get if isinstance(clip, dict) else lambda k, d=None: getattr(clip, k, d)
    if not get("ENABLED", False):
        return 0.0
    if get("CLIP_TYPE", "value") == "full_model":
        return float(get("CLIP_VALUE")) if get("CLIP_VALUE", 0.0) > 0.0 else 0.0
    # build.py:206-223 falls back to detectron2's per-parameter clipping for "value" / "norm": not a flat-buffer operation
    raise capi.Yb200Error("SOLVER.CLIP_GRADIENTS.CLIP_TYPE %r is not implemented by the flat optimizer (only 'full_model'); "
                          "disable clipping or use full_model" % (get("CLIP_TYPE", "value"),))


def _build(cfg, model, kind, **kw):
    eng = _flat_buffers(model)
    overrides = {}
    for d in _solver(cfg, "LR_MULTIPLIER_OVERWRITE", None) or []:
        overrides.update(d)  # build.py:226-231 _merge_dict
    if hasattr(eng, "flat_buffers"):
        bufs = eng.flat_buffers()
    else:
        bufs = [(eng.flat_param, eng.flat_grad, eng.param_layout, getattr(eng, "norm_param_names", None))]
    if len(bufs) > 1 and _clip_norm(cfg) > 0:
        raise capi.Yb200Error("full-model gradient clipping over several flat buffers is not implemented")
    opts = []
    for flat_param, flat_grad, layout, norm_names in bufs:
        segs = param_segments(layout, flat_param.numel(), _solver(cfg, "WEIGHT_DECAY", 1e-4), _solver(cfg, "WEIGHT_DECAY_NORM", None),
                              _solver(cfg, "WEIGHT_DECAY_BIAS", None), bias_lr_factor=_solver(cfg, "BIAS_LR_FACTOR", 1.0),
                              lr_multipliers_overwrite=overrides, norm_param_names=norm_names)
        opts.append(FlatOptimizer(flat_param, flat_grad, segs, _solver(cfg, "BASE_LR", 0.001), kind, clip_norm=_clip_norm(cfg), **kw))
    return opts[0] if len(opts) == 1 else FlatOptimizerGroup(opts)


def sgd(cfg, model):
    """build.py:234-245"""
    return _build(cfg, model, "sgd", momentum=_solver(cfg, "MOMENTUM", 0.9), nesterov=_solver(cfg, "NESTEROV", False))


def adamw(cfg, model):
    """build.py:248-256"""
    return _build(cfg, model, "adamw")


sgd_mt, adamw_mt = sgd, adamw  # build.py:259-288: the multi-tensor variants are what this file always is

_MAPPER = {"sgd": sgd, "adamw": adamw, "sgd_mt": sgd_mt, "adamw_mt": adamw_mt}


def build_optimizer_mapper(cfg, model):
    """build.py:291-300"""
    name = str(_solver(cfg, "OPTIMIZER", "sgd")).lower()
    if name not in _MAPPER:
        raise KeyError(f"No object named '{name}' found in 'D2GO_OPTIM_MAPPER' registry!")
    return _MAPPER[name](cfg, model)
