"""Seeded synthetic head outputs and ground truth for the ERD loss path.

Shapes and distributions follow SURVEY.md section 8(d): student class logits
~ N(-4.6, 1) (the head's ``bias_prob=0.01`` init, reference gfl_head.py:115-123),
box-distribution logits ~ N(0, 1); GT boxes in the ``demo_mm_inputs`` convention
(reference mmdet/testing/_utils.py:66-75) with continuous coordinates.  Everything is
generated on the CPU from explicit generators so the oracle and the CUDA path see
identical bits.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

STRIDES = (8, 16, 32, 64, 128)


def level_shapes(pad_h: int, pad_w: int, strides: Sequence[int] = STRIDES) -> List[Tuple[int, int]]:
    h, w = math.ceil(pad_h / strides[0]), math.ceil(pad_w / strides[0])
    out = [(h, w)]
    for _ in strides[1:]:
        h, w = (h + 1) // 2, (w + 1) // 2
        out.append((h, w))
    return out


@dataclass
class Batch:
    t_cls: List[torch.Tensor]
    t_box: List[torch.Tensor]
    s_cls: List[torch.Tensor]
    s_box: List[torch.Tensor]
    gt_bboxes: List[torch.Tensor]
    gt_labels: List[torch.Tensor]
    pad_shapes: List[Tuple[int, int]]
    img_shapes: List[Tuple[int, int]]
    ori: int
    num_classes: int
    reg_max: int
    canvas: Tuple[int, int]
    shapes: List[Tuple[int, int]] = field(default_factory=list)

    @property
    def num_imgs(self) -> int:
        return self.s_cls[0].size(0)

    @property
    def anchors_per_image(self) -> int:
        return sum(h * w for h, w in self.shapes)

    def to(self, device) -> 'Batch':
        mv = lambda xs: [x.to(device) for x in xs]
        return Batch(mv(self.t_cls), mv(self.t_box), mv(self.s_cls), mv(self.s_box),
                     mv(self.gt_bboxes), mv(self.gt_labels), list(self.pad_shapes),
                     list(self.img_shapes), self.ori, self.num_classes, self.reg_max,
                     self.canvas, list(self.shapes))


def make_gt(rng: np.random.RandomState, num: int, img_h: int, img_w: int, num_new: int,
            size_pow: float = 1.0):
    """demo_mm_inputs-style boxes: centre and size ~ U(0,1) of the image, clipped.
    ``size_pow`` > 1 skews sizes small (U**pow) so fine pyramid levels get positives."""
    cx, cy, bw, bh = rng.rand(4, num)
    if size_pow != 1.0:
        bw, bh = bw ** size_pow, bh ** size_pow
    x1 = ((cx * img_w) - (img_w * bw / 2)).clip(0, img_w)
    y1 = ((cy * img_h) - (img_h * bh / 2)).clip(0, img_h)
    x2 = ((cx * img_w) + (img_w * bw / 2)).clip(0, img_w)
    y2 = ((cy * img_h) + (img_h * bh / 2)).clip(0, img_h)
    boxes = np.stack([x1, y1, x2, y2], 1).astype(np.float32)
    labels = rng.randint(0, num_new, size=num).astype(np.int64)
    return torch.from_numpy(boxes), torch.from_numpy(labels)


def make_batch(num_imgs: int = 2, img_hw: Tuple[int, int] = (800, 1333), ori: int = 40,
               num_classes: int = 80, reg_max: int = 16, seed: int = 1234,
               num_gt: Optional[object] = None, mode: str = 'gaussian',
               pad_shapes: Optional[Sequence[Tuple[int, int]]] = None,
               img_shapes: Optional[Sequence[Tuple[int, int]]] = None,
               pad_divisor: int = 32, gt_size_pow: float = 1.0) -> Batch:
    """``num_gt``: None -> U{1..9} per image; int -> fixed; list -> per image;
    (lo, hi) tuple -> U{lo..hi}.
    ``mode='trained'`` plants confident teacher responses (about 1 % of anchors at
    N(+1,1) class logits and a +6 peak on one box bin per side) so teacher boxes
    overlap and NMS has work to do.  ``pad_shapes`` gives per-image padded shapes
    (<= the batch canvas) to exercise the valid-flag logic."""
    img_h, img_w = img_hw
    ch = int(math.ceil(img_h / pad_divisor) * pad_divisor)
    cw = int(math.ceil(img_w / pad_divisor) * pad_divisor)
    shapes = level_shapes(ch, cw)
    g = torch.Generator().manual_seed(seed)
    r = reg_max + 1

    def randn(*shape, mean=0.0):
        t = torch.randn(*shape, generator=g, dtype=torch.float32)
        return t + mean if mean else t

    s_cls = [randn(num_imgs, num_classes, h, w, mean=-4.6) for h, w in shapes]
    s_box = [randn(num_imgs, 4 * r, h, w) for h, w in shapes]
    t_cls = [randn(num_imgs, ori, h, w, mean=-4.6) for h, w in shapes]
    t_box = [randn(num_imgs, 4 * r, h, w) for h, w in shapes]
    if mode == 'trained':
        for lv, (h, w) in enumerate(shapes):
            # iid confident anchors: ~1 % get one class logit lifted to ~N(+1,1)
            hot = torch.rand(num_imgs, 1, h, w, generator=g) < 0.01
            cls_pick = torch.randint(0, ori, (num_imgs, 1, h, w), generator=g)
            t_cls[lv] = t_cls[lv] + torch.zeros_like(t_cls[lv]).scatter_(1, cls_pick, 5.6) * hot
            tb = t_box[lv].view(num_imgs, 4, r, h, w)
            bins = torch.randint(r // 2, r, (num_imgs, 4, 1, h, w), generator=g)
            tb += torch.zeros_like(tb).scatter_(2, bins, 6.0) * (torch.rand(num_imgs, 1, 1, h, w, generator=g) < 0.01)
            # object-like blobs: a 5x5 patch shares one class and one peaked box distribution,
            # so neighbouring teacher boxes overlap and class-aware NMS suppresses most of them
            if h >= 8 and w >= 8:
                for i in range(num_imgs):
                    for _ in range(12):
                        cy = int(torch.randint(2, h - 2, (1,), generator=g))
                        cx = int(torch.randint(2, w - 2, (1,), generator=g))
                        c = int(torch.randint(0, ori, (1,), generator=g))
                        bsel = torch.randint(r // 2, r, (4,), generator=g)
                        t_cls[lv][i, c, cy - 2:cy + 3, cx - 2:cx + 3] += 5.6
                        for k in range(4):
                            tb[i, k, int(bsel[k]), cy - 2:cy + 3, cx - 2:cx + 3] += 6.0
    elif mode != 'gaussian':
        raise ValueError(mode)

    rng = np.random.RandomState(seed)
    if img_shapes is None:
        img_shapes = [(img_h, img_w)] * num_imgs
    if pad_shapes is None:
        pad_shapes = [(ch, cw)] * num_imgs
    gt_b, gt_l = [], []
    for i in range(num_imgs):
        if num_gt is None:
            k = int(rng.randint(1, 10))
        elif isinstance(num_gt, int):
            k = num_gt
        elif isinstance(num_gt, list):
            k = int(num_gt[i])
        else:
            k = int(rng.randint(num_gt[0], num_gt[1] + 1))
        b, l = make_gt(rng, k, img_shapes[i][0], img_shapes[i][1], num_classes - ori, gt_size_pow)
        gt_b.append(b)
        gt_l.append(l)
    return Batch(t_cls, t_box, s_cls, s_box, gt_b, gt_l, [tuple(p) for p in pad_shapes],
                 [tuple(p) for p in img_shapes], ori, num_classes, reg_max, (ch, cw), shapes)
