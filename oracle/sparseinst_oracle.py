"""TEST INFRASTRUCTURE -- CPU restatement (plain torch fp32) of the reference's SparseInst IAM decoder forward.

Pinned by tests/golden/sparseinst.npz, produced by oracle/gen_golden_sparseinst.py from the UNMODIFIED reference classes
(yolov7/modeling/transcoders/decoder_sparseinst.py:27-169: InstanceBranch, MaskBranch, BaseIAMDecoder; fvcore / detectron2 helpers stubbed).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it.
"""
import torch
import torch.nn.functional as F


def coordinates(x):
    """BaseIAMDecoder.compute_coordinates, decoder_sparseinst.py:118-127: (x_loc, y_loc) in [-1, 1], prepended to the features"""
    h, w = x.shape[2], x.shape[3]
    y_loc = torch.linspace(-1, 1, h)
    x_loc = torch.linspace(-1, 1, w)
    y_loc, x_loc = torch.meshgrid(y_loc, x_loc, indexing="ij")
    y_loc = y_loc.expand(x.shape[0], 1, -1, -1)
    x_loc = x_loc.expand(x.shape[0], 1, -1, -1)
    return torch.cat([x_loc, y_loc], 1).to(x)


def _stack(x, sd, prefix, n):
    for i in range(n):  # _make_stack_3x3_convs :18-24: Conv2d(3x3, pad 1) + ReLU; Sequential indices 0, 2, 4, ...
        x = F.relu(F.conv2d(x, sd[f"{prefix}{2 * i}.weight"], sd[f"{prefix}{2 * i}.bias"], padding=1))
    return x


def instance_branch(features, sd, prefix="inst_branch.", num_convs=4):
    """InstanceBranch.forward, decoder_sparseinst.py:62-81"""
    f = _stack(features, sd, prefix + "inst_convs.", num_convs)
    iam = F.conv2d(f, sd[prefix + "iam_conv.weight"], sd[prefix + "iam_conv.bias"], padding=1)          # :66
    prob = iam.sigmoid()                                                                                   # :67
    b, n = prob.shape[:2]
    c = f.shape[1]
    prob = prob.view(b, n, -1)
    inst = torch.bmm(prob, f.view(b, c, -1).permute(0, 2, 1))                                              # :74
    inst = inst / prob.sum(-1).clamp(min=1e-6)[:, :, None]                                                 # :75-76
    logits = F.linear(inst, sd[prefix + "cls_score.weight"], sd[prefix + "cls_score.bias"])              # :78
    kernel = F.linear(inst, sd[prefix + "mask_kernel.weight"], sd[prefix + "mask_kernel.bias"])          # :79
    scores = F.linear(inst, sd[prefix + "objectness.weight"], sd[prefix + "objectness.bias"])            # :80
    return logits, kernel, scores, iam


def group_instance_branch(features, sd, prefix="inst_branch.", num_convs=4, groups=4):
    """GroupInstanceBranch.forward, decoder_sparseinst.py:212-242: grouped IAM convolution (N masks per group), aggregation, the four groups of
    one instance concatenated along the channels, fc + ReLU, then the heads"""
    f = _stack(features, sd, prefix + "inst_convs.", num_convs)
    iam = F.conv2d(f, sd[prefix + "iam_conv.weight"], sd[prefix + "iam_conv.bias"], padding=1, groups=groups)   # :215
    prob = iam.sigmoid()
    b, n = prob.shape[:2]
    c = f.shape[1]
    prob = prob.view(b, n, -1)
    inst = torch.bmm(prob, f.view(b, c, -1).permute(0, 2, 1))                                                # :224
    inst = inst / prob.sum(-1).clamp(min=1e-6, max=1e5)[:, :, None]                                         # :225-227
    d4 = n // 4                                                                                              # :230
    inst = inst.reshape(b, 4, d4, -1).transpose(1, 2).reshape(b, d4, -1)                                     # :231-235
    inst = F.relu(F.linear(inst, sd[prefix + "fc.weight"], sd[prefix + "fc.bias"]))                        # :237
    logits = F.linear(inst, sd[prefix + "cls_score.weight"], sd[prefix + "cls_score.bias"])
    kernel = F.linear(inst, sd[prefix + "mask_kernel.weight"], sd[prefix + "mask_kernel.bias"])
    scores = F.linear(inst, sd[prefix + "objectness.weight"], sd[prefix + "objectness.bias"])
    return logits, kernel, scores, iam


def mask_branch(features, sd, prefix="mask_branch.", num_convs=4):
    """MaskBranch.forward, decoder_sparseinst.py:101-104"""
    f = _stack(features, sd, prefix + "mask_convs.", num_convs)
    return F.conv2d(f, sd[prefix + "projection.weight"], sd[prefix + "projection.bias"])


def decoder_forward(features, sd, scale_factor=2.0, num_convs=4, groups=0):
    """BaseIAMDecoder.forward, decoder_sparseinst.py:130-169 (groups > 0: GroupIAMDecoder :245-250, the same forward with the group branch)"""
    x = torch.cat([coordinates(features), features], 1)                                                    # :131-132
    if groups:
        logits, kernel, scores, iam = group_instance_branch(x, sd, num_convs=num_convs, groups=groups)
    else:
        logits, kernel, scores, iam = instance_branch(x, sd, num_convs=num_convs)
    mf = mask_branch(x, sd, num_convs=num_convs)
    b, c, h, w = mf.shape
    masks = torch.bmm(kernel, mf.view(b, c, h * w)).view(b, kernel.shape[1], h, w)                         # :143-146
    masks = F.interpolate(masks, scale_factor=scale_factor, mode="bilinear", align_corners=False)         # :148-153
    return {"pred_logits": logits, "pred_masks": masks, "pred_scores": scores, "pred_kernel": kernel, "iam": iam, "masks_lowres": torch.bmm(kernel, mf.view(b, c, h * w)).view(b, -1, h, w)}


def decoder_state_dict(seed=0, in_channels=256, dim=256, num_masks=100, kernel_dim=128, num_classes=80, num_convs=4, trained_like=True, groups=0):
    g = torch.Generator().manual_seed(seed)

    def rn(*s, std):
        return torch.randn(*s, generator=g) * std

    sd = {}
    cin = in_channels + 2
    for br, d in (("inst_branch.inst_convs.", dim), ("mask_branch.mask_convs.", dim)):
        c = cin
        for i in range(num_convs):
            sd[f"{br}{2 * i}.weight"] = rn(d, c, 3, 3, std=(2.0 / (9 * c)) ** 0.5)
            sd[f"{br}{2 * i}.bias"] = rn(d, std=0.05) if trained_like else torch.zeros(d)
            c = d
    nm_all, hd = (num_masks * groups, dim * groups) if groups else (num_masks, dim)  # GroupInstanceBranch :182-193
    sd["inst_branch.iam_conv.weight"] = rn(nm_all, dim // groups if groups else dim, 3, 3, std=0.02 if trained_like else 0.01)
    sd["inst_branch.iam_conv.bias"] = torch.full((nm_all,), -2.0 if trained_like else -4.595) + (rn(nm_all, std=0.5) if trained_like else 0)
    if groups:
        sd["inst_branch.fc.weight"], sd["inst_branch.fc.bias"] = rn(hd, hd, std=(1.0 / hd) ** 0.5), rn(hd, std=0.05)
    sd["inst_branch.cls_score.weight"], sd["inst_branch.cls_score.bias"] = rn(num_classes, hd, std=0.05), rn(num_classes, std=0.5) - 2.0
    sd["inst_branch.mask_kernel.weight"], sd["inst_branch.mask_kernel.bias"] = rn(kernel_dim, hd, std=0.05), rn(kernel_dim, std=0.1)
    sd["inst_branch.objectness.weight"], sd["inst_branch.objectness.bias"] = rn(1, hd, std=0.05), rn(1, std=0.1)
    sd["mask_branch.projection.weight"], sd["mask_branch.projection.bias"] = rn(kernel_dim, dim, 1, 1, std=(2.0 / dim) ** 0.5), rn(kernel_dim, std=0.05)
    return sd
