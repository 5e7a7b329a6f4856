"""Synthetic COCO-shaped inputs for benchmarks, profiling tools and tests.  Host-side data generation only (seeded torch
CPU generators): nothing here is on the measured path.  The oracle modules import these generators too, so every arm of bench.py and every
golden fixture draws from the same distributions."""
import math

import torch


def synthetic_batch(batch, size=640, seed=0, max_gt=20, max_boxes=100, empty_every=0):
    """uint8 images U{0..255}; labels [B, max_boxes, 5] = (cls, cx, cy, w, h) with 1..max_gt boxes per image (cls U{0..79},
    centre U(0.1, 0.9) * size, w / h log-uniform(16, 0.6 * size)), clipped to the image, zero rows elsewhere."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randint(0, 256, (batch, 3, size, size), generator=g, dtype=torch.uint8)
    labels = torch.zeros(batch, max_boxes, 5)
    for b in range(batch):
        if empty_every and (b % empty_every) == empty_every - 1:
            continue
        k = int(torch.randint(1, max_gt + 1, (1,), generator=g))
        cxy = (torch.rand(k, 2, generator=g) * 0.8 + 0.1) * size
        lo, hi = math.log(16.0), math.log(0.6 * size)
        wh = torch.exp(torch.rand(k, 2, generator=g) * (hi - lo) + lo)
        x1y1 = (cxy - wh / 2).clamp(0, size)
        x2y2 = (cxy + wh / 2).clamp(0, size)
        labels[b, :k, 0] = torch.randint(0, 80, (k,), generator=g).float()
        labels[b, :k, 1:3] = (x1y1 + x2y2) / 2
        labels[b, :k, 3:5] = (x2y2 - x1y1).clamp(min=2.0)
    return images, labels


def clustered_predictions(batch, anchors, num_classes, seed):
    """NMS stress set: per image 30 ground-truth boxes x jittered copies (uniform-random boxes almost never overlap and make NMS trivial);
    [batch, anchors, 5 + num_classes] = (cx, cy, w, h, obj, class probabilities)."""
    g = torch.Generator().manual_seed(seed)
    pred = torch.zeros(batch, anchors, 5 + num_classes)
    for b in range(batch):
        gt = torch.cat([torch.rand(30, 2, generator=g) * 500 + 70, torch.exp(torch.rand(30, 2, generator=g) * 2.5 + 2.5)], 1)
        cls = torch.randint(0, num_classes, (30,), generator=g)
        idx = torch.arange(anchors) % 30
        box = gt[idx] * (1 + torch.randn(anchors, 4, generator=g) * 0.1)
        box[:, 2:] = box[:, 2:].abs() + 1
        pred[b, :, :4] = box
        u = torch.rand(anchors, 2, generator=g)
        pred[b, :, 4] = u[:, 0] ** 0.6 * 0.9 + 0.05
        c = cls[idx].clone()
        flip = torch.rand(anchors, generator=g) < 0.1
        c[flip] = torch.randint(0, num_classes, (int(flip.sum()),), generator=g)
        pred[b, :, 5:] = torch.rand(anchors, num_classes, generator=g) * 0.05
        pred[b, torch.arange(anchors), 5 + c] = u[:, 1] ** 0.5 * 0.9 + 0.08
    return pred


def synthetic_images(batch, size, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (batch, 3, size, size), generator=g, dtype=torch.uint8)
