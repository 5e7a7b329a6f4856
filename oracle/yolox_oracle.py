"""TEST INFRASTRUCTURE -- CPU oracle: a functional fp32 restatement of the reference's YOLOX hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this file; the
product path (yolov7_d2_b200/) never does.  Every function cites the reference code it restates (paths relative to
lucasjinreal/yolov7_d2 @ 7805129).  Parity is pinned: tests/test_oracle_golden.py checks this file against
tests/golden/*.npz, which oracle/gen_golden.py produced by executing the reference's own files (oracle/ref_shim.py).

Everything is driven by a `state_dict` with the reference's parameter names (backbone.*, neck.*, head.*), so the
same weights feed the reference, this oracle and the CUDA path.  All math is torch CPU fp32 (autograd gives the
backward oracle).  The number of bottlenecks per CSP stage is read off the state_dict keys.
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-3       # yolox.py:85-90 (_init_model)
BN_MOMENTUM = 0.03  # yolox.py:85-90
STRIDES = (8, 16, 32)  # yolox_head.py:29


# --------------------------------------------------------------------------------------------------------
# building blocks  (yolov7/modeling/backbone/layers/wrappers.py)
# --------------------------------------------------------------------------------------------------------
def silu(x):
    return x * torch.sigmoid(x)


# Optional instrumentation (tests only): TRACE collects every BaseConv output by parameter prefix; EMULATE_STORAGE rounds
# the stored tensors exactly where the CUDA engine stores 16-bit values (pre-BN conv output: fp16; activation and
# residual sum: bf16), giving an oracle that differs from the engine only by fp32 summation order -- the yardstick for
# "how far may a faithful 16-bit implementation be from the fp32 reference".
TRACE = None
EMULATE_STORAGE = False


def _qz(t):
    return t.to(torch.float16).float() if EMULATE_STORAGE else t


def _q(t):
    return t.to(torch.bfloat16).float() if EMULATE_STORAGE else t


def base_conv(x, sd, prefix, stride=1, training=True):
    """BaseConv = Conv2d(bias=False, pad=(k-1)//2) -> BatchNorm2d -> SiLU   (wrappers.py:60-80)"""
    w = sd[prefix + ".conv.weight"]
    z = _qz(F.conv2d(x, w, None, stride, (w.shape[-1] - 1) // 2))
    z = F.batch_norm(z, sd[prefix + ".bn.running_mean"], sd[prefix + ".bn.running_var"], sd[prefix + ".bn.weight"],
                     sd[prefix + ".bn.bias"], training, BN_MOMENTUM, BN_EPS)
    if training and (prefix + ".bn.num_batches_tracked") in sd:
        sd[prefix + ".bn.num_batches_tracked"] += 1
    a = _q(silu(z))
    if TRACE is not None:
        TRACE[prefix] = a.detach()
    return a


def focus(x):
    """space-to-depth, channel order (top-left, bottom-left, top-right, bottom-right)   (wrappers.py:210-220)"""
    return torch.cat((x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]), 1)


def bottleneck(x, sd, prefix, shortcut, training):
    """1x1 -> 3x3, residual when shortcut and cin == cout   (wrappers.py:105-123)"""
    y = base_conv(base_conv(x, sd, prefix + ".conv1", 1, training), sd, prefix + ".conv2", 1, training)
    if shortcut and y.shape[1] == x.shape[1]:
        y = _q(y + x)
        if TRACE is not None:
            TRACE[prefix + ".conv2"] = y.detach()  # what the engine stores for this op: the residual sum
    return y


def csp_layer(x, sd, prefix, shortcut, training):
    """C3: conv1 -> n bottlenecks, conv2, cat, conv3   (wrappers.py:165-199)"""
    a = base_conv(x, sd, prefix + ".conv1", 1, training)
    b = base_conv(x, sd, prefix + ".conv2", 1, training)
    i = 0
    while f"{prefix}.m.{i}.conv1.conv.weight" in sd:
        a = bottleneck(a, sd, f"{prefix}.m.{i}", shortcut, training)
        i += 1
    return base_conv(torch.cat((a, b), 1), sd, prefix + ".conv3", 1, training)


def spp_bottleneck(x, sd, prefix, training):
    """1x1, max-pool 5/9/13 (stride 1, pad k//2), cat, 1x1   (wrappers.py:142-162)"""
    x = base_conv(x, sd, prefix + ".conv1", 1, training)
    x = torch.cat([x] + [F.max_pool2d(x, k, 1, k // 2) for k in (5, 9, 13)], 1)
    return base_conv(x, sd, prefix + ".conv2", 1, training)


# --------------------------------------------------------------------------------------------------------
# backbone / neck / head   (darknetx.py:103-177, yolo_pafpn.py:79-114, yolox_head.py:151-245)
# --------------------------------------------------------------------------------------------------------
def csp_darknet(x, sd, training=True, P="backbone."):
    x = base_conv(focus(x), sd, P + "stem.conv", 1, training)
    feats = {}
    for name in ("dark2", "dark3", "dark4"):
        x = base_conv(x, sd, f"{P}{name}.0", 2, training)
        x = csp_layer(x, sd, f"{P}{name}.1", True, training)
        feats[name] = x
    x = base_conv(x, sd, P + "dark5.0", 2, training)
    x = spp_bottleneck(x, sd, P + "dark5.1", training)
    feats["dark5"] = csp_layer(x, sd, P + "dark5.2", False, training)
    return feats


def pafpn(feats, sd, training=True, P="neck."):
    x2, x1, x0 = feats["dark3"], feats["dark4"], feats["dark5"]
    up = lambda t: F.interpolate(t, scale_factor=2, mode="nearest")
    fpn_out0 = base_conv(x0, sd, P + "lateral_conv0", 1, training)
    f_out0 = csp_layer(torch.cat([up(fpn_out0), x1], 1), sd, P + "C3_p4", False, training)
    fpn_out1 = base_conv(f_out0, sd, P + "reduce_conv1", 1, training)
    pan_out2 = csp_layer(torch.cat([up(fpn_out1), x2], 1), sd, P + "C3_p3", False, training)
    p_out1 = base_conv(pan_out2, sd, P + "bu_conv2", 2, training)
    pan_out1 = csp_layer(torch.cat([p_out1, fpn_out1], 1), sd, P + "C3_n3", False, training)
    p_out0 = base_conv(pan_out1, sd, P + "bu_conv1", 2, training)
    pan_out0 = csp_layer(torch.cat([p_out0, fpn_out0], 1), sd, P + "C3_n4", False, training)
    return pan_out2, pan_out1, pan_out0


def head_raw(fpn_outs, sd, training=True, P="head."):
    """per level [B, 5+C, H, W] = cat(reg(4), obj(1), cls(C)) raw conv outputs   (yolox_head.py:160-175)"""
    outs = []
    for k, x in enumerate(fpn_outs):
        x = base_conv(x, sd, f"{P}stems.{k}", 1, training)
        c = base_conv(base_conv(x, sd, f"{P}cls_convs.{k}.0", 1, training), sd, f"{P}cls_convs.{k}.1", 1, training)
        r = base_conv(base_conv(x, sd, f"{P}reg_convs.{k}.0", 1, training), sd, f"{P}reg_convs.{k}.1", 1, training)
        cls = F.conv2d(c, sd[f"{P}cls_preds.{k}.weight"], sd[f"{P}cls_preds.{k}.bias"])
        reg = F.conv2d(r, sd[f"{P}reg_preds.{k}.weight"], sd[f"{P}reg_preds.{k}.bias"])
        obj = F.conv2d(r, sd[f"{P}obj_preds.{k}.weight"], sd[f"{P}obj_preds.{k}.bias"])
        outs.append(torch.cat([reg, obj, cls], 1))
    return outs


def anchor_grid(hw_list, strides=STRIDES, dtype=torch.float32, device=None):
    """x_shifts, y_shifts, expanded_strides, each [A]   (yolox_head.py:226-245, 295-300)"""
    xs, ys, ss = [], [], []
    for (h, w), s in zip(hw_list, strides):
        yv, xv = torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device), indexing="ij")
        xs.append(xv.reshape(-1).to(dtype))
        ys.append(yv.reshape(-1).to(dtype))
        ss.append(torch.full((h * w,), float(s), dtype=dtype, device=device))
    return torch.cat(xs), torch.cat(ys), torch.cat(ss)


def decode_train(raw_levels, strides=STRIDES):
    """[B, A, 5+C]: xy = (xy + grid) * s, wh = exp(wh) * s, obj / cls logits untouched   (yolox_head.py:226-245)"""
    outs = []
    for o, s in zip(raw_levels, strides):
        b, ch, h, w = o.shape
        o = o.permute(0, 2, 3, 1).reshape(b, h * w, ch)
        gx, gy, _ = anchor_grid([(h, w)], [s], o.dtype, o.device)
        grid = torch.stack((gx, gy), 1).unsqueeze(0)
        outs.append(torch.cat([(o[..., :2] + grid) * s, torch.exp(o[..., 2:4]) * s, o[..., 4:]], -1))
    return torch.cat(outs, 1)


def decode_eval(raw_levels, strides=STRIDES):
    """inference branch: sigmoid on obj / cls, then the same box decode   (yolox_head.py:197-224, 247-272)"""
    outs = []
    for o, s in zip(raw_levels, strides):
        b, ch, h, w = o.shape
        o = torch.cat([o[:, :4], o[:, 4:].sigmoid()], 1).permute(0, 2, 3, 1).reshape(b, h * w, ch)
        gx, gy, _ = anchor_grid([(h, w)], [s], o.dtype, o.device)
        grid = torch.stack((gx, gy), 1).unsqueeze(0)
        outs.append(torch.cat([(o[..., :2] + grid) * s, torch.exp(o[..., 2:4]) * s, o[..., 4:]], -1))
    return torch.cat(outs, 1)


# --------------------------------------------------------------------------------------------------------
# boxes / losses   (yolov7/utils/boxes.py)
# --------------------------------------------------------------------------------------------------------
def bboxes_iou_cxcywh(a, b):
    """pairwise IoU [N,M] of cxcywh boxes, no epsilon   (boxes.py:57-81, xyxy=False branch)"""
    tl = torch.max(a[:, None, :2] - a[:, None, 2:] / 2, b[:, :2] - b[:, 2:] / 2)
    br = torch.min(a[:, None, :2] + a[:, None, 2:] / 2, b[:, :2] + b[:, 2:] / 2)
    area_a = a[:, 2] * a[:, 3]
    area_b = b[:, 2] * b[:, 3]
    en = (tl < br).to(tl.dtype).prod(2)
    inter = (br - tl).prod(2) * en
    return inter / (area_a[:, None] + area_b - inter)


def iou_loss(pred, target, loss_type="iou"):
    """IOUloss(reduction='none'): 1 - iou^2, or 1 - clamp(giou)   (boxes.py:125-168)"""
    tl = torch.max(pred[:, :2] - pred[:, 2:] / 2, target[:, :2] - target[:, 2:] / 2)
    br = torch.min(pred[:, :2] + pred[:, 2:] / 2, target[:, :2] + target[:, 2:] / 2)
    area_p = pred[:, 2] * pred[:, 3]
    area_g = target[:, 2] * target[:, 3]
    en = (tl < br).to(tl.dtype).prod(1)
    inter = (br - tl).prod(1) * en
    iou = inter / (area_p + area_g - inter + 1e-16)
    if loss_type == "iou":
        return 1 - iou ** 2
    c_tl = torch.min(pred[:, :2] - pred[:, 2:] / 2, target[:, :2] - target[:, 2:] / 2)
    c_br = torch.max(pred[:, :2] + pred[:, 2:] / 2, target[:, :2] + target[:, 2:] / 2)
    area_c = (c_br - c_tl).prod(1)
    giou = iou - (area_c - inter) / area_c.clamp(1e-16)
    return 1 - giou.clamp(min=-1.0, max=1.0)


def iou_loss_v6(pred, target, iou_type="ciou", eps=1e-7):
    """IOUlossV6(box_format='xywh', reduction='none') for giou / diou / ciou; pred, target are [F,4] cxcywh
    (the reference is called with box1 already transposed: boxes.py:682-752)."""
    p, t = pred.T, target.T
    b1x1, b1x2, b1y1, b1y2 = p[0] - p[2] / 2, p[0] + p[2] / 2, p[1] - p[3] / 2, p[1] + p[3] / 2
    b2x1, b2x2, b2y1, b2y2 = t[0] - t[2] / 2, t[0] + t[2] / 2, t[1] - t[3] / 2, t[1] + t[3] / 2
    inter = (torch.min(b1x2, b2x2) - torch.max(b1x1, b2x1)).clamp(0) * (torch.min(b1y2, b2y2) - torch.max(b1y1, b2y1)).clamp(0)
    w1, h1 = b1x2 - b1x1, b1y2 - b1y1 + eps
    w2, h2 = b2x2 - b2x1, b2y2 - b2y1 + eps
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    cw = torch.max(b1x2, b2x2) - torch.min(b1x1, b2x1)
    ch = torch.max(b1y2, b2y2) - torch.min(b1y1, b2y1)
    if iou_type == "giou":
        c_area = cw * ch + eps
        iou = iou - (c_area - union) / c_area
    elif iou_type in ("diou", "ciou"):
        c2 = cw ** 2 + ch ** 2 + eps
        rho2 = ((b2x1 + b2x2 - b1x1 - b1x2) ** 2 + (b2y1 + b2y2 - b1y1 - b1y2) ** 2) / 4
        if iou_type == "diou":
            iou = iou - rho2 / c2
        else:
            v = (4 / math.pi ** 2) * (torch.atan(w2 / h2) - torch.atan(w1 / h1)) ** 2
            with torch.no_grad():
                alpha = v / (v - iou + (1 + eps))
            iou = iou - (rho2 / c2 + v * alpha)
    else:
        raise ValueError(iou_type)
    return 1.0 - iou


# --------------------------------------------------------------------------------------------------------
# SimOTA   (yolox_head.py:450-669)
# --------------------------------------------------------------------------------------------------------
def in_boxes_info(gt, x_shifts, y_shifts, strides, radius=2.5):
    """candidate anchors (inside any gt box OR any 2.5-stride centre square) and the per-(gt,candidate)
    in-box-AND-in-centre matrix; all tests are strict `> 0`   (yolox_head.py:549-633)"""
    xc = (x_shifts * strides + 0.5 * strides)[None]
    yc = (y_shifts * strides + 0.5 * strides)[None]
    l, r = (gt[:, 0] - 0.5 * gt[:, 2])[:, None], (gt[:, 0] + 0.5 * gt[:, 2])[:, None]
    t, b = (gt[:, 1] - 0.5 * gt[:, 3])[:, None], (gt[:, 1] + 0.5 * gt[:, 3])[:, None]
    in_box = torch.stack([xc - l, yc - t, r - xc, b - yc], 2).min(-1).values > 0.0
    rs = radius * strides[None]
    cl, cr = gt[:, 0:1] - rs, gt[:, 0:1] + rs
    ct, cb = gt[:, 1:2] - rs, gt[:, 1:2] + rs
    in_ctr = torch.stack([xc - cl, yc - ct, cr - xc, cb - yc], 2).min(-1).values > 0.0
    cand = (in_box.sum(0) > 0) | (in_ctr.sum(0) > 0)
    return cand, in_box[:, cand] & in_ctr[:, cand]


def simota_assign(gt_boxes, gt_classes, boxes, cls_logits, obj_logits, x_shifts, y_shifts, strides):
    """One image.  Returns (fg_mask[A] bool, matched_gt[F] int64, matched_cls[F], matched_iou[F]) with F = fg_mask.sum().
    cost = BCE(sqrt(sig(cls)*sig(obj)), onehot).sum(-1) + 3*(-log(iou+1e-8)) + 1e5*(not in box-and-centre);
    k_g = clamp(int(sum of the 10 largest IoUs of gt g), 1); each gt takes its k_g cheapest candidates; a candidate
    claimed by several gts goes to the cheapest   (yolox_head.py:479-532, 635-669)."""
    num_classes = cls_logits.shape[1]
    cand, in_both = in_boxes_info(gt_boxes, x_shifts, y_shifts, strides)
    cb, cc, co = boxes[cand], cls_logits[cand], obj_logits[cand]
    m = cb.shape[0]
    ious = bboxes_iou_cxcywh(gt_boxes, cb)
    onehot = F.one_hot(gt_classes.to(torch.int64), num_classes).float()
    prob = (cc.float().sigmoid() * co.float().sigmoid()[:, None]).sqrt()
    cls_cost = F.binary_cross_entropy(prob[None].expand(gt_boxes.shape[0], m, num_classes),
                                      onehot[:, None].expand(-1, m, -1), reduction="none").sum(-1)
    cost = cls_cost + 3.0 * (-torch.log(ious + 1e-8)) + 100000.0 * (~in_both)

    matching = torch.zeros_like(cost)
    topk_ious, _ = torch.topk(ious, min(10, m), dim=1)
    ks = torch.clamp(topk_ious.sum(1).int(), min=1)
    for g in range(gt_boxes.shape[0]):
        _, pos = torch.topk(cost[g], k=int(ks[g]), largest=False)
        matching[g][pos] = 1.0
    multi = matching.sum(0) > 1
    if multi.any():
        arg = cost[:, multi].argmin(0)
        matching[:, multi] = 0.0
        matching[arg, multi] = 1.0
    fg_in = matching.sum(0) > 0
    fg_mask = cand.clone()
    fg_mask[cand] = fg_in
    matched_gt = matching[:, fg_in].argmax(0)
    return fg_mask, matched_gt, gt_classes[matched_gt], (matching * ious).sum(0)[fg_in]


def l1_target(gt, stride, x_shifts, y_shifts, eps=1e-8):
    """get_l1_target (yolox_head.py:443-448): the matched gt box in the RAW parametrisation of the regression outputs"""
    return torch.stack([gt[:, 0] / stride - x_shifts, gt[:, 1] / stride - y_shifts, torch.log(gt[:, 2] / stride + eps), torch.log(gt[:, 3] / stride + eps)], 1)


def yolox_losses(outputs, labels, x_shifts, y_shifts, strides, num_classes=80, loss_type="iou", return_assign=False, origin_preds=None):
    """YOLOXHead.get_losses   (yolox_head.py:274-441).
    outputs [B,A,5+C] (decoded boxes + raw logits), labels [B,G,5] = (cls,cx,cy,w,h) zero padded.
    Returns (total, 5*iou, obj, cls, num_fg/num_gt).  origin_preds [B,A,4] (the raw regression outputs, `use_l1`, :186-195, :389-429) adds the
    L1 term to the total and makes the result (total, 5*iou, obj, cls, l1, num_fg/num_gt)."""
    bsz, num_anchors = outputs.shape[:2]
    nlabel = (labels.sum(2) > 0).sum(1)
    fg_masks, cls_t, reg_t, assigns, l1_t = [], [], [], [], []
    num_fg, num_gts = 0.0, 0.0
    for b in range(bsz):
        g = int(nlabel[b])
        num_gts += g
        if g == 0:
            fg = torch.zeros(num_anchors, dtype=torch.bool, device=outputs.device)
            cls_t.append(outputs.new_zeros((0, num_classes)))
            reg_t.append(outputs.new_zeros((0, 4)))
            assigns.append((fg, torch.zeros(0, dtype=torch.int64, device=outputs.device), outputs.new_zeros(0), outputs.new_zeros(0)))
        else:
            gtb, gtc = labels[b, :g, 1:5], labels[b, :g, 0]
            with torch.no_grad():
                fg, mgt, mcls, miou = simota_assign(gtb, gtc, outputs[b, :, :4], outputs[b, :, 5:], outputs[b, :, 4],
                                                    x_shifts, y_shifts, strides)
            num_fg += int(fg.sum())
            cls_t.append(F.one_hot(mcls.to(torch.int64), num_classes) * miou[:, None])
            reg_t.append(gtb[mgt])
            assigns.append((fg, mgt, mcls, miou))
            if origin_preds is not None:
                l1_t.append(l1_target(gtb[mgt], strides[fg], x_shifts[fg], y_shifts[fg]))
        fg_masks.append(fg)
    fg_all = torch.cat(fg_masks)
    cls_t, reg_t = torch.cat(cls_t), torch.cat(reg_t)
    obj_t = fg_all.to(outputs.dtype)[:, None]
    num_fg = max(num_fg, 1)
    l_iou = iou_loss(outputs[..., :4].reshape(-1, 4)[fg_all], reg_t, loss_type).sum() / num_fg
    l_obj = F.binary_cross_entropy_with_logits(outputs[..., 4].reshape(-1, 1), obj_t, reduction="none").sum() / num_fg
    l_cls = F.binary_cross_entropy_with_logits(outputs[..., 5:].reshape(-1, num_classes)[fg_all], cls_t, reduction="none").sum() / num_fg
    total = 5.0 * l_iou + l_obj + l_cls
    if origin_preds is not None:
        tgt = torch.cat(l1_t) if l1_t else outputs.new_zeros((0, 4))
        l_l1 = (origin_preds.reshape(-1, 4)[fg_all] - tgt).abs().sum() / num_fg
        res = (total + l_l1, 5.0 * l_iou, l_obj, l_cls, l_l1, num_fg / max(num_gts, 1))
    else:
        res = (total, 5.0 * l_iou, l_obj, l_cls, num_fg / max(num_gts, 1))
    return res + (assigns,) if return_assign else res


# --------------------------------------------------------------------------------------------------------
# whole model
# --------------------------------------------------------------------------------------------------------
def yolox_forward_train(images, labels, sd, num_classes=80):
    """images [B,3,H,W] float (0..255, un-normalised: yolox.py:98), labels [B,G,5].  YOLOX.forward training branch
    (yolox.py:192-209): returns the 4 loss scalars + num_fg ratio + the decoded head output."""
    fpn = pafpn(csp_darknet(images, sd, True), sd, True)
    raw = head_raw(fpn, sd, True)
    outputs = decode_train(raw)
    xs, ys, ss = anchor_grid([o.shape[-2:] for o in raw], device=outputs.device)
    return yolox_losses(outputs, labels, xs, ys, ss, num_classes) + (outputs,)


def yolox_forward_eval(images, sd):
    fpn = pafpn(csp_darknet(images, sd, False), sd, False)
    return decode_eval(head_raw(fpn, sd, False))


# --------------------------------------------------------------------------------------------------------
# post-processing   (boxes.py:171-210 + torchvision.ops.batched_nms, per-class "vanilla" semantics)
# --------------------------------------------------------------------------------------------------------
def nms_greedy(boxes, scores, thr):
    """torchvision.ops.nms semantics: stable descending score order; box j is suppressed by a kept box i when
    inter / (area_i + area_j - inter) > thr (fp32, xyxy, no +1)."""
    order = torch.sort(scores, descending=True, stable=True).indices
    b = boxes[order]
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    n = b.shape[0]
    suppressed = torch.zeros(n, dtype=torch.bool)
    keep = []
    for i in range(n):
        if suppressed[i]:
            continue
        keep.append(i)
        if i + 1 < n:
            xx1 = torch.maximum(b[i, 0], b[i + 1:, 0])
            yy1 = torch.maximum(b[i, 1], b[i + 1:, 1])
            xx2 = torch.minimum(b[i, 2], b[i + 1:, 2])
            yy2 = torch.minimum(b[i, 3], b[i + 1:, 3])
            inter = (xx2 - xx1).clamp(min=0) * (yy2 - yy1).clamp(min=0)
            ovr = inter / (area[i] + area[i + 1:] - inter)
            suppressed[i + 1:] |= ovr > thr
    return order[torch.tensor(keep, dtype=torch.int64)]


def batched_nms_vanilla(boxes, scores, classes, thr, tie_order="stable"):
    """per-class NMS on un-offset coordinates, survivors ordered by descending score (torchvision
    `_batched_nms_vanilla`; SURVEY.md par.0.3 explains why this is the canonical semantics).
    tie_order: how EQUAL scores are ordered in the result --
      "stable"         lower index first: what torchvision's `nms` returns (stable sort in both its CPU and CUDA kernels), hence what the
                       reference produces on CUDA (coordinate-trick strategy for < 25 000 boxes) and on the CPU for <= 1000 candidates;
      "torch_cpu_sort" the literal last line of `_batched_nms_vanilla`, `scores[keep].sort(descending=True)`: torch's UNSTABLE CPU sort, whose
                       permutation of ties is an artefact of its introsort (pinned by tests/golden/nms_ties.npz, `big` case)."""
    keep = torch.zeros(scores.shape[0], dtype=torch.bool)
    for c in torch.unique(classes):
        idx = torch.where(classes == c)[0]
        keep[idx[nms_greedy(boxes[idx], scores[idx], thr)]] = True
    k = torch.where(keep)[0]
    if tie_order == "torch_cpu_sort":
        return k[scores[k].sort(descending=True)[1]]
    assert tie_order == "stable", tie_order
    return k[torch.sort(scores[k], descending=True, stable=True).indices]


def postprocess(prediction, num_classes, conf_thre, nms_thre, tie_order="stable"):
    """boxes.py:171-210 (without the in-place mutation): per image [n,7] = (x1,y1,x2,y2,obj,cls_conf,cls) or None."""
    pred = prediction.clone()
    pred[..., 0] = prediction[..., 0] - prediction[..., 2] / 2
    pred[..., 1] = prediction[..., 1] - prediction[..., 3] / 2
    pred[..., 2] = prediction[..., 0] + prediction[..., 2] / 2
    pred[..., 3] = prediction[..., 1] + prediction[..., 3] / 2
    out = []
    for p in pred:
        conf, cls = p[:, 5:5 + num_classes].max(1, keepdim=True)
        mask = (p[:, 4] * conf.squeeze(1)) >= conf_thre
        det = torch.cat((p[:, :5], conf, cls.float()), 1)[mask]
        if det.shape[0] == 0:
            out.append(None)
            continue
        out.append(det[batched_nms_vanilla(det[:, :4], det[:, 4] * det[:, 5], det[:, 6], nms_thre, tie_order)])
    return out


def fuse_conv_bn(w, gamma, beta, mean, var, eps=BN_EPS):
    """eval-mode folding: w' = w * gamma/sqrt(var+eps), b' = beta - mean*gamma/sqrt(var+eps)   (utils/checkpoint.py:11-43)"""
    s = gamma / torch.sqrt(var + eps)
    return w * s.view(-1, 1, 1, 1), beta - mean * s


def preprocess(images_u8, pad_value=114.0, divisibility=32):
    """list of uint8 [3,h,w] -> float [B,3,H,W] padded bottom/right to a multiple of 32 with 114, no mean/std
    (yolox.py:95-102 + detectron2 ImageList.from_tensors)."""
    hmax = max(i.shape[1] for i in images_u8)
    wmax = max(i.shape[2] for i in images_u8)
    hp = (hmax + divisibility - 1) // divisibility * divisibility
    wp = (wmax + divisibility - 1) // divisibility * divisibility
    out = torch.full((len(images_u8), 3, hp, wp), pad_value, dtype=torch.float32)
    for k, im in enumerate(images_u8):
        out[k, :, :im.shape[1], :im.shape[2]] = im.float()
    return out


# --------------------------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------------------------
def yolox_state_dict(seed=0, width=0.5, depth=0.33, num_classes=80, prior_prob=1e-2):
    """Random-init YOLOX state_dict with the reference's names / shapes and its default initialisers
    (nn.Conv2d kaiming-uniform a=sqrt(5); BN gamma=1, beta=0; prediction biases -log((1-p)/p): yolox_head.py:140-149).
    Weights are rounded to bf16-representable values so that the CUDA path (bf16 operands) and this fp32 oracle
    consume identical numbers."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(prefix, cin, cout, k, bias=False):
        fan_in = cin * k * k
        bound = 1.0 / math.sqrt(fan_in)  # kaiming_uniform_(a=sqrt(5)) bound = sqrt(6/((1+5)*fan_in))
        w = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
        sd[prefix + ".weight"] = w.to(torch.bfloat16).float()
        if bias:
            sd[prefix + ".bias"] = ((torch.rand(cout, generator=g) * 2 - 1) * bound).to(torch.bfloat16).float()

    def bconv(prefix, cin, cout, k):
        conv(prefix + ".conv", cin, cout, k)
        sd[prefix + ".bn.weight"] = torch.ones(cout)
        sd[prefix + ".bn.bias"] = torch.zeros(cout)
        sd[prefix + ".bn.running_mean"] = torch.zeros(cout)
        sd[prefix + ".bn.running_var"] = torch.ones(cout)
        sd[prefix + ".bn.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)

    def csp(prefix, cin, cout, n):
        h = int(cout * 0.5)
        bconv(prefix + ".conv1", cin, h, 1)
        bconv(prefix + ".conv2", cin, h, 1)
        bconv(prefix + ".conv3", 2 * h, cout, 1)
        for i in range(n):
            bconv(f"{prefix}.m.{i}.conv1", h, h, 1)
            bconv(f"{prefix}.m.{i}.conv2", h, h, 3)

    bc = int(width * 64)
    bd = max(round(depth * 3), 1)
    bconv("backbone.stem.conv", 12, bc, 3)
    bconv("backbone.dark2.0", bc, bc * 2, 3); csp("backbone.dark2.1", bc * 2, bc * 2, bd)
    bconv("backbone.dark3.0", bc * 2, bc * 4, 3); csp("backbone.dark3.1", bc * 4, bc * 4, bd * 3)
    bconv("backbone.dark4.0", bc * 4, bc * 8, 3); csp("backbone.dark4.1", bc * 8, bc * 8, bd * 3)
    bconv("backbone.dark5.0", bc * 8, bc * 16, 3)
    bconv("backbone.dark5.1.conv1", bc * 16, bc * 8, 1); bconv("backbone.dark5.1.conv2", bc * 32, bc * 16, 1)
    csp("backbone.dark5.2", bc * 16, bc * 16, bd)
    c0, c1, c2 = int(256 * width), int(512 * width), int(1024 * width)
    n = round(3 * depth)
    bconv("neck.lateral_conv0", c2, c1, 1); csp("neck.C3_p4", 2 * c1, c1, n)
    bconv("neck.reduce_conv1", c1, c0, 1); csp("neck.C3_p3", 2 * c0, c0, n)
    bconv("neck.bu_conv2", c0, c0, 3); csp("neck.C3_n3", 2 * c0, c1, n)
    bconv("neck.bu_conv1", c1, c1, 3); csp("neck.C3_n4", 2 * c1, c2, n)
    hc = int(256 * width)
    for k, cin in enumerate((c0, c1, c2)):
        bconv(f"head.stems.{k}", cin, hc, 1)
        for br in ("cls_convs", "reg_convs"):
            bconv(f"head.{br}.{k}.0", hc, hc, 3)
            bconv(f"head.{br}.{k}.1", hc, hc, 3)
        conv(f"head.cls_preds.{k}", hc, num_classes, 1, bias=True)
        conv(f"head.reg_preds.{k}", hc, 4, 1, bias=True)
        conv(f"head.obj_preds.{k}", hc, 1, 1, bias=True)
        prior = -math.log((1 - prior_prob) / prior_prob)
        sd[f"head.cls_preds.{k}.bias"] = torch.full((num_classes,), prior)
        sd[f"head.obj_preds.{k}.bias"] = torch.full((1,), prior)
    return sd


from yolov7_d2_b200.synth import synthetic_batch  # noqa: E402,F401  (shared generator: same inputs for every arm and every fixture)
