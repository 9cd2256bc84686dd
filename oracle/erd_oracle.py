"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the GFL+ERD dense-head loss path.

A torch-CPU restatement of the reference algorithm (Hi-FT/ERD, an MMDetection
3.0.0 fork).  It is the *checker* for the CUDA path in ``erd_b200``: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it.  The product never does.

Every function cites the reference ``file:line`` it follows (paths relative to
the reference root).  fp32 operation order is kept where an integer output
depends on it (IoU, centre distance, thresholds), so on CPU this file
reproduces the reference bit-for-bit; ``tests/test_oracle_vs_reference.py``
checks that against the real reference loaded by path
(``oracle/ref_by_path.py``) and ``tests/golden/`` holds vectors generated from
the real reference by ``oracle/make_golden.py``.

PARITY UNPINNED for ``batched_nms`` only: ``mmcv.ops.batched_nms`` (mmcv
>=2.0.0rc4,<2.1.0, ``requirements/mminstall.txt:1``) is third-party and absent
from the reference tree and the reference's tests hold no NMS vector; it is
restated here from mmcv 2.0.x published semantics and cross-checked against
``torchvision.ops.nms``.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

EPS32 = float(torch.finfo(torch.float32).eps)
STRIDES = (8, 16, 32, 64, 128)
INF = 100000000


# --------------------------------------------------------------------------- geometry
def level_shapes(pad_h: int, pad_w: int, strides: Sequence[int] = STRIDES) -> List[Tuple[int, int]]:
    """Feature-map sizes a R50-FPN (start_level=1, extra convs stride 2) emits for a
    padded input: each level is ceil(previous / 2) (SURVEY Appendix B shapes)."""
    h, w = math.ceil(pad_h / strides[0]), math.ceil(pad_w / strides[0])
    out = [(h, w)]
    for _ in strides[1:]:
        h, w = (h + 1) // 2, (w + 1) // 2
        out.append((h, w))
    return out


def level_anchors(h: int, w: int, stride: int, scale: float = 8.0) -> Tensor:
    """Square anchors of one level, row-major (y*W+x).
    anchor_generator.py:161-205 (base anchor, ratio 1, centre_offset 0) and
    :266-301 (shift grid): (x*s - 4s, y*s - 4s, x*s + 4s, y*s + 4s)."""
    half = 0.5 * stride * scale
    xs = torch.arange(w, dtype=torch.float32) * stride
    ys = torch.arange(h, dtype=torch.float32) * stride
    yy, xx = torch.meshgrid(ys, xs, indexing='ij')
    xx, yy = xx.reshape(-1), yy.reshape(-1)
    return torch.stack([xx - half, yy - half, xx + half, yy + half], dim=1)


def level_valid_flags(h: int, w: int, stride: int, pad_h: int, pad_w: int) -> Tensor:
    """anchor_generator.py:415-476: x < min(ceil(pad_w/s), W), y < min(ceil(pad_h/s), H)."""
    vh = min(int(math.ceil(pad_h / stride)), h)
    vw = min(int(math.ceil(pad_w / stride)), w)
    fy = torch.arange(h) < vh
    fx = torch.arange(w) < vw
    return (fy[:, None] & fx[None, :]).reshape(-1)


# --------------------------------------------------------------------------- IoU family
def pairwise_iou(a: Tensor, b: Tensor, eps: float = 1e-6) -> Tensor:
    """bbox_overlaps.py:151-153,170-193 (mode='iou', is_aligned=False)."""
    if a.size(0) * b.size(0) == 0:
        return a.new_zeros((a.size(0), b.size(0)))
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = area_a[:, None] + area_b[None, :] - inter
    union = torch.max(union, union.new_tensor([eps]))
    return inter / union


def aligned_iou(a: Tensor, b: Tensor, giou: bool = False, eps: float = 1e-6) -> Tensor:
    """bbox_overlaps.py:151-169,189-199 (is_aligned=True; 'iou' or 'giou')."""
    if a.size(0) == 0:
        return a.new_zeros((0,))
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, :2], b[:, :2])
    rb = torch.min(a[:, 2:], b[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    union = area_a + area_b - inter
    e = union.new_tensor([eps])
    union = torch.max(union, e)
    iou = inter / union
    if not giou:
        return iou
    elt = torch.min(a[:, :2], b[:, :2])
    erb = torch.max(a[:, 2:], b[:, 2:])
    ewh = (erb - elt).clamp(min=0)
    earea = torch.max(ewh[:, 0] * ewh[:, 1], e)
    return iou - (earea - union) / earea


# --------------------------------------------------------------------------- ATSS
def atss_assign(priors: Tensor, num_level_priors: Sequence[int], gt_bboxes: Tensor,
                gt_labels: Tensor, topk: int = 9, report: Optional[dict] = None):
    """atss_assigner.py:74-254 with alpha=None, ignore_iof_thr=-1.

    Returns (assigned_gt_inds int64 (A,), max_overlaps (A,), assigned_labels (A,)).
    ``report`` (optional dict) receives tie / margin diagnostics used by the parity
    tests to tell an arbitrary-tie or sub-ulp threshold flip from a real mismatch
    (SURVEY Appendix C 9, 9b)."""
    num_gt, num_priors = gt_bboxes.size(0), priors.size(0)
    overlaps = pairwise_iou(priors, gt_bboxes)                                   # :138
    assigned = overlaps.new_zeros((num_priors,), dtype=torch.long)               # :162
    if num_gt == 0 or num_priors == 0:                                           # :166-176
        return (assigned, overlaps.new_zeros((num_priors,)),
                overlaps.new_full((num_priors,), -1, dtype=torch.long))
    gcx = (gt_bboxes[:, 0] + gt_bboxes[:, 2]) / 2.0                              # :25-26
    gcy = (gt_bboxes[:, 1] + gt_bboxes[:, 3]) / 2.0
    pcx = (priors[:, 0] + priors[:, 2]) / 2.0                                    # :29-30
    pcy = (priors[:, 1] + priors[:, 3]) / 2.0
    ppts = torch.stack((pcx, pcy), dim=1)
    gpts = torch.stack((gcx, gcy), dim=1)
    dist = (ppts[:, None, :] - gpts[None, :, :]).pow(2).sum(-1).sqrt()           # :33-34
    cand, start = [], 0
    for n_l in num_level_priors:                                                 # :193-202
        k = min(topk, n_l)
        d_l = dist[start:start + n_l]
        _, idx = d_l.topk(k, dim=0, largest=False)
        if report is not None and n_l > k:
            kp1 = d_l.topk(k + 1, dim=0, largest=False)[0]
            ties = (kp1[k - 1] == kp1[k]).nonzero().flatten().tolist() if k > 0 else []
            report.setdefault('topk_boundary_ties', []).extend(ties)
        cand.append(idx + start)
        start += n_l
    cand = torch.cat(cand, dim=0)                                                # :203
    cand_ov = overlaps[cand, torch.arange(num_gt)]                               # :207
    thr = cand_ov.mean(0) + cand_ov.std(0)                                       # :208-210
    is_pos = cand_ov >= thr[None, :]                                             # :212
    if report is not None:
        report['min_thr_margin'] = float((cand_ov - thr[None, :]).abs().min())
    ccx, ccy = pcx[cand], pcy[cand]                                              # :217-230
    l_ = ccx - gt_bboxes[:, 0]
    t_ = ccy - gt_bboxes[:, 1]
    r_ = gt_bboxes[:, 2] - ccx
    b_ = gt_bboxes[:, 3] - ccy
    inside = torch.stack([l_, t_, r_, b_], dim=1).min(dim=1)[0] > 0.01           # :231
    is_pos = is_pos & inside                                                     # :233
    table = torch.full((num_gt, num_priors), -float(INF), dtype=overlaps.dtype)  # :237-241
    gcol = torch.arange(num_gt)[None, :].expand_as(cand)
    table[gcol[is_pos], cand[is_pos]] = overlaps[cand[is_pos], gcol[is_pos]]
    max_ov, arg = table.t().max(dim=1)                                           # :243
    hit = max_ov != -INF
    assigned[hit] = arg[hit] + 1                                                 # :244-245
    labels = assigned.new_full((num_priors,), -1)                                # :247-252
    pos = assigned > 0
    labels[pos] = gt_labels[assigned[pos] - 1]
    return assigned, max_ov, labels


def image_targets(anchors: Tensor, valid: Tensor, num_level_anchors: Sequence[int],
                  gt_bboxes: Tensor, gt_labels: Tensor, num_classes: int,
                  report: Optional[dict] = None):
    """gfl_head.py:562-679 (_get_targets_single with allowed_border=-1, pos_weight=-1,
    PseudoSampler pseudo_sampler.py:26-60, unmap misc.py:222-232).

    Returns dict(anchors, labels, label_weights, bbox_targets, gt_inds, num_pos) over
    ALL anchors; ``gt_inds`` is -1 on invalid (padded-region) anchors."""
    if not bool(valid.any()):
        raise ValueError('There is no valid anchor inside the image boundary.')   # :613-617
    a = anchors[valid]
    starts = [0]
    for n in num_level_anchors:
        starts.append(starts[-1] + n)
    inside_counts = [int(valid[starts[i]:starts[i + 1]].sum()) for i in range(len(num_level_anchors))]
    gt_inds, _, _ = atss_assign(a, inside_counts, gt_bboxes, gt_labels, report=report)
    pos = (gt_inds > 0).nonzero().flatten()
    n_valid = a.size(0)
    tgt = torch.zeros_like(a)
    lab = torch.full((n_valid,), num_classes, dtype=torch.long)
    lw = torch.zeros(n_valid, dtype=torch.float32)
    if pos.numel() > 0:
        tgt[pos] = gt_bboxes[gt_inds[pos] - 1]                                    # sampling_result.py:86-116
        lab[pos] = gt_labels[gt_inds[pos] - 1]
        lw[pos] = 1.0
    lw[gt_inds == 0] = 1.0
    total = anchors.size(0)

    def unmap(x, fill):
        out = x.new_full((total,) + tuple(x.shape[1:]), fill)
        out[valid] = x
        return out
    return dict(anchors=unmap(a, 0), labels=unmap(lab, num_classes), label_weights=unmap(lw, 0),
                bbox_targets=unmap(tgt, 0), gt_inds=unmap(gt_inds, -1),
                num_pos=int(pos.numel()))


# --------------------------------------------------------------------------- loss pieces
def integral(box_logits: Tensor, reg_max: int = 16) -> Tensor:
    """gfl_head_increment_erd.py:40-54: softmax over reg_max+1 bins, expectation."""
    p = F.softmax(box_logits.reshape(-1, reg_max + 1), dim=1)
    proj = torch.linspace(0, reg_max, reg_max + 1).type_as(p)
    return F.linear(p, proj).reshape(-1, 4)


def points_to_box(pts: Tensor, d: Tensor) -> Tensor:
    """transforms.py:169-174 (distance2bbox, max_shape=None)."""
    return torch.stack([pts[:, 0] - d[:, 0], pts[:, 1] - d[:, 1],
                        pts[:, 0] + d[:, 2], pts[:, 1] + d[:, 3]], -1)


def box_to_distances(pts: Tensor, box: Tensor, reg_max: int) -> Tensor:
    """transforms.py:221-230 with max_dis=reg_max, eps=0.1 (coder :28-53)."""
    hi = reg_max - 0.1
    return torch.stack([(pts[:, 0] - box[:, 0]).clamp(min=0, max=hi),
                        (pts[:, 1] - box[:, 1]).clamp(min=0, max=hi),
                        (box[:, 2] - pts[:, 0]).clamp(min=0, max=hi),
                        (box[:, 3] - pts[:, 1]).clamp(min=0, max=hi)], -1)


def qfl_elementwise(pred: Tensor, label: Tensor, score: Tensor, beta: float = 2.0) -> Tensor:
    """gfocal_loss.py:12-53, per-anchor (summed over classes)."""
    sig = pred.sigmoid()
    loss = F.binary_cross_entropy_with_logits(pred, sig.new_zeros(pred.shape), reduction='none') * sig.pow(beta)
    pos = ((label >= 0) & (label < pred.size(1))).nonzero().squeeze(1)
    pl = label[pos].long()
    sf = score[pos] - sig[pos, pl]
    loss[pos, pl] = F.binary_cross_entropy_with_logits(pred[pos, pl], score[pos], reduction='none') * sf.abs().pow(beta)
    return loss.sum(dim=1)


def dfl_elementwise(pred: Tensor, label: Tensor) -> Tensor:
    """gfocal_loss.py:143-165."""
    lo = label.long()
    hi = lo + 1
    return (F.cross_entropy(pred, lo, reduction='none') * (hi.float() - label)
            + F.cross_entropy(pred, hi, reduction='none') * (label - lo.float()))


def kd_kl_elementwise(pred: Tensor, soft: Tensor, T: float) -> Tensor:
    """kd_loss.py:12-37."""
    target = F.softmax(soft / T, dim=1).detach()
    return F.kl_div(F.log_softmax(pred / T, dim=1), target, reduction='none').mean(1) * (T * T)


def reduce_with_avg(loss: Tensor, weight: Tensor, avg_factor: float) -> Tensor:
    """losses/utils.py:30-65 (reduction='mean' with avg_factor)."""
    return (loss * weight).sum() / (avg_factor + EPS32)


def level_loss(anchors: Tensor, cls_score: Tensor, bbox_pred: Tensor, labels: Tensor,
               label_weights: Tensor, bbox_targets: Tensor, stride: int, num_classes: int,
               ori: int, avg_factor: float, reg_max: int = 16):
    """gfl_head_increment_erd.py:225-322 for one pyramid level.

    cls_score (N,C,H,W), bbox_pred (N,4R,H,W); labels etc. (N,A_l[,4]).
    Returns loss_cls, loss_bbox, loss_dfl (each with fixed inner avg factors applied),
    sum of weight_targets, and the per-anchor IoU score."""
    cn = num_classes - ori
    anchors = anchors.reshape(-1, 4)
    cls = cls_score[:, ori:].permute(0, 2, 3, 1).reshape(-1, cn)                  # :260-261
    box = bbox_pred.permute(0, 2, 3, 1).reshape(-1, 4 * (reg_max + 1))            # :263-264
    tgt = bbox_targets.reshape(-1, 4)
    lab = labels.reshape(-1).clone()
    lw = label_weights.reshape(-1)
    lab[lab == num_classes] = cn                                                  # :270-271
    pos = ((lab >= 0) & (lab < cn)).nonzero().squeeze(1)                          # :273-274
    score = lw.new_zeros(lab.shape, dtype=cls.dtype)   # fp32 on the reference path; fp64 for the accuracy study
    if pos.numel() > 0:
        pa = anchors[pos]
        ctr = torch.stack([(pa[:, 0] + pa[:, 2]) / 2, (pa[:, 1] + pa[:, 3]) / 2], -1) / stride  # gfl_head.py:232-243
        w = cls.detach().sigmoid().max(dim=1)[0][pos]                             # :283-284
        pb = box[pos]
        dec = points_to_box(ctr, integral(pb, reg_max))                           # :285-287
        t = tgt[pos] / stride                                                     # :288
        score[pos] = aligned_iou(dec.detach(), t)                                 # :289-292
        corners = box_to_distances(ctr, t, reg_max).reshape(-1)                   # :294-296
        loss_bbox = 2.0 * reduce_with_avg(1 - aligned_iou(dec, t, giou=True), w, 1.0)   # :299-303, iou_loss.py:520-527
        loss_dfl = 0.25 * reduce_with_avg(dfl_elementwise(pb.reshape(-1, reg_max + 1), corners),
                                          w[:, None].expand(-1, 4).reshape(-1), 4.0)    # :306-310
        wsum = w.sum()
    else:
        loss_bbox = box.sum() * 0                                                 # :311-314
        loss_dfl = box.sum() * 0
        wsum = box.new_tensor(0).sum()
    loss_cls = 1.0 * reduce_with_avg(qfl_elementwise(cls, lab, score), lw, avg_factor)   # :317-320
    return loss_cls, loss_bbox, loss_dfl, wsum, score


# --------------------------------------------------------------------------- ERS
def flatten_levels(levels: Sequence[Tensor]) -> Tensor:
    """(N,C,H_l,W_l) x L -> (N, A, C): gfl_increment_erd.py:183-193 / ERD head :412-434."""
    n = levels[0].size(0)
    return torch.cat([t.permute(0, 2, 3, 1).reshape(n, -1, t.size(1)) for t in levels], dim=1)


def ers_select_single(cls_a: Tensor, box_a: Tensor, report: Optional[dict] = None):
    """gfl_increment_erd.py:143-163: rows above mean + 2*std (unbiased) of the
    per-anchor max sigmoid score / max raw box logit."""
    m = cls_a.sigmoid().max(dim=-1)[0]
    thr_c = m.mean() + 2 * m.std()
    cls_inds = (m > thr_c).nonzero(as_tuple=False).squeeze(1)
    u = box_a.max(dim=-1)[0]
    thr_b = u.mean() + 2 * u.std()
    box_inds = (u > thr_b).nonzero(as_tuple=False).squeeze(1)
    if report is not None:
        report.setdefault('cls_thr', []).append(float(thr_c))
        report.setdefault('box_thr', []).append(float(thr_b))
        report.setdefault('cls_margin', []).append(float((m - thr_c).abs().min()))
        report.setdefault('box_margin', []).append(float((u - thr_b).abs().min()))
    return cls_inds, box_inds


def sel_pos(t_cls: Sequence[Tensor], t_box: Sequence[Tensor], report: Optional[dict] = None):
    """gfl_increment_erd.py:165-200. Returns (cls_inds list[N], box_inds list[N])."""
    ca, ba = flatten_levels(t_cls), flatten_levels(t_box)
    out = [ers_select_single(ca[i], ba[i], report) for i in range(ca.size(0))]
    return [o[0] for o in out], [o[1] for o in out]


# --------------------------------------------------------------------------- NMS (third-party, restated)
def nms(boxes: Tensor, scores: Tensor, iou_threshold: float) -> Tensor:
    """Greedy IoU NMS with mmcv 2.0.x ``nms`` (offset=0) CPU semantics: visit boxes by
    descending score, area (x2-x1)*(y2-y1), suppress when
    inter / (area_i + area_j - inter) > thr; returns kept indices in visiting order.
    Ties in score are visited lowest index first (mmcv's sort is unstable, so any order
    is 'reference behaviour'; the tests flag score ties)."""
    n = boxes.size(0)
    if n == 0:
        return torch.zeros(0, dtype=torch.long)
    order = torch.sort(scores, descending=True, stable=True)[1]
    import numpy as np
    b = boxes[order].numpy().astype(np.float32)
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    alive = np.ones(n, dtype=bool)
    keep = []
    for i in range(n):
        if not alive[i]:
            continue
        keep.append(i)
        xx1 = np.maximum(x1[i], x1[i + 1:])
        yy1 = np.maximum(y1[i], y1[i + 1:])
        xx2 = np.minimum(x2[i], x2[i + 1:])
        yy2 = np.minimum(y2[i], y2[i + 1:])
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide='ignore', invalid='ignore'):
            ovr = inter / (areas[i] + areas[i + 1:] - inter)
        alive[i + 1:] &= ~(ovr > np.float32(iou_threshold))
    return order[torch.as_tensor(keep, dtype=torch.long)]


def batched_nms(boxes: Tensor, scores: Tensor, idxs: Tensor, nms_cfg: dict, class_agnostic: bool = False):
    """mmcv 2.0.x ``batched_nms`` semantics: class-aware via the coordinate-offset trick
    ``boxes + idxs * (boxes.max() + 1)``, one NMS call below ``split_thr`` (10000) boxes.
    Returns (dets (M,5), keep (M,)) with keep in descending-score order.
    Call site: gfl_head_increment_erd.py:198-202 (iou_threshold=0.005)."""
    cfg = dict(nms_cfg)
    class_agnostic = cfg.pop('class_agnostic', class_agnostic)
    cfg.pop('type', None)
    cfg.pop('split_thr', None)
    thr = cfg['iou_threshold']
    if boxes.numel() == 0:
        return torch.cat([boxes, scores[:, None]], -1), torch.zeros(0, dtype=torch.long)
    if class_agnostic:
        b = boxes
    else:
        mx = boxes.max()
        off = idxs.to(boxes) * (mx + torch.tensor(1).to(boxes))
        b = boxes + off[:, None]
    keep = nms(b.detach(), scores.detach(), thr)
    return torch.cat([boxes[keep], scores[keep][:, None]], -1), keep


# --------------------------------------------------------------------------- distillation
def distill_image(anchors: Tensor, s_cls_old: Tensor, s_box: Tensor, cls_inds: Tensor,
                  box_inds: Tensor, t_cls: Tensor, t_box: Tensor, dist_loss_weight: float,
                  ori: int, reg_max: int = 16, T: float = 10.0, report: Optional[dict] = None):
    """gfl_head_increment_erd.py:142-223 for one image.
    anchors (A,4) [zeros on padded anchors], s_cls_old (A,ori), s_box (A,4R), teacher same."""
    a = s_cls_old[cls_inds] - t_cls[cls_inds]                                     # :181-186
    loss_cls = dist_loss_weight * a.pow(2).float().mean()                         # :324-332
    ctr = torch.stack([(anchors[:, 0] + anchors[:, 2]) / 2, (anchors[:, 1] + anchors[:, 3]) / 2], -1)
    tb = points_to_box(ctr, integral(t_box, reg_max))                             # :189-192 (no *stride)
    conf, ids = t_cls.sigmoid().max(dim=-1)                                       # :194-195
    _, keep = batched_nms(tb[box_inds], conf[box_inds], ids[box_inds], dict(iou_threshold=0.005))  # :198-202
    rows = box_inds[keep]
    sp = s_box[rows].reshape(-1, reg_max + 1)                                     # :204-215
    tp = t_box[rows].reshape(-1, reg_max + 1)
    w = s_cls_old.reshape(-1, ori)[box_inds].detach().sigmoid().max(dim=1)[0][keep.reshape(-1)]  # :217-218
    loss_box = dist_loss_weight * 0.25 * reduce_with_avg(
        kd_kl_elementwise(sp, tp, T), w[:, None].expand(-1, 4).reshape(-1), 4.0)  # :219-221, kd_loss.py:60-95
    if report is not None:
        report.setdefault('keep', []).append(keep)
        sc = conf[box_inds]
        report.setdefault('score_ties', []).append(int(sc.numel() - torch.unique(sc).numel()))
    return loss_cls, loss_box


# --------------------------------------------------------------------------- the whole path
def loss_by_feat(t_cls: Sequence[Tensor], t_box: Sequence[Tensor], s_cls: Sequence[Tensor],
                 s_box: Sequence[Tensor], cls_inds: Sequence[Tensor], box_inds: Sequence[Tensor],
                 ori: int, dist_loss_weight: float, gt_bboxes: Sequence[Tensor],
                 gt_labels: Sequence[Tensor], pad_shapes: Sequence[Tuple[int, int]],
                 num_classes: int = 80, reg_max: int = 16, strides: Sequence[int] = STRIDES,
                 reduce_mean=None, report: Optional[dict] = None):
    """gfl_head_increment_erd.py:334-454.  ``reduce_mean``: the reference's
    ``mmdet.utils.reduce_mean`` (dist_utils.py:59-65: ``t.clone().div_(world).all_reduce(SUM)``)
    as a callable on 0-dim float tensors; None = not distributed (passthrough, :61-62).  It is
    applied at the reference's two call sites (:390-391 and :406-407), which is how the
    multi-rank tests feed this rank the world-averaged normalisers.
    Returns the reference's loss dict."""
    if reduce_mean is None:
        reduce_mean = lambda t: t
    n = s_cls[0].size(0)
    sizes = [tuple(t.shape[-2:]) for t in s_cls]
    assert len(sizes) == len(strides)                                             # :374
    anchors_l = [level_anchors(h, w, s) for (h, w), s in zip(sizes, strides)]     # anchor_head.py:164-199
    n_level = [a.size(0) for a in anchors_l]
    all_anchors = torch.cat(anchors_l)
    per_img = []
    for i in range(n):
        valid = torch.cat([level_valid_flags(h, w, s, pad_shapes[i][0], pad_shapes[i][1])
                           for (h, w), s in zip(sizes, strides)])
        rep_i = {} if report is not None else None
        per_img.append(image_targets(all_anchors, valid, n_level, gt_bboxes[i], gt_labels[i],
                                     num_classes, report=rep_i))
        if report is not None:
            report.setdefault('atss', []).append(rep_i)
            report.setdefault('gt_inds', []).append(per_img[-1]['gt_inds'])
    avg1 = float(sum(max(t['num_pos'], 1) for t in per_img))                      # gfl_head.py:548-549, sampling_result.py:96-100
    avg1_local = avg1
    avg1 = reduce_mean(torch.tensor(avg1, dtype=torch.float)).item()              # :390-391

    def by_level(key):                                                            # misc.py:427-440
        stacked = torch.stack([t[key] for t in per_img], 0)
        out, s0 = [], 0
        for nl in n_level:
            out.append(stacked[:, s0:s0 + nl])
            s0 += nl
        return out
    anc_l, lab_l, lw_l, tgt_l = by_level('anchors'), by_level('labels'), by_level('label_weights'), by_level('bbox_targets')
    l_cls, l_bbox, l_dfl, wsums, scores = [], [], [], [], []
    for lv, s in enumerate(strides):                                              # :393-404
        a, b, c, wsum, sc = level_loss(anc_l[lv], s_cls[lv], s_box[lv], lab_l[lv], lw_l[lv], tgt_l[lv],
                                       s, num_classes, ori, avg1, reg_max)
        l_cls.append(a); l_bbox.append(b); l_dfl.append(c); wsums.append(wsum); scores.append(sc)
    wsum_local = sum(wsums)
    avg2 = reduce_mean(wsum_local.detach().clone()).clamp_(min=1).item()          # :406-407
    l_bbox = [x / avg2 for x in l_bbox]                                           # :408-409
    l_dfl = [x / avg2 for x in l_dfl]
    if report is not None:
        report['avg_factors'] = (avg1, avg2)
        report['avg_local'] = (avg1_local, float(wsum_local))
        report['scores'] = scores

    anc = torch.cat(anc_l, dim=1)                                                 # :412
    sb = flatten_levels(s_box)                                                    # :413-416
    tc = flatten_levels([t[:, :ori] for t in t_cls])                              # :420-424
    tb = flatten_levels(t_box)                                                    # :426-429
    sc_old = flatten_levels([t[:, :ori] for t in s_cls])                          # :431-434
    d_cls, d_box = [], []
    for i in range(n):                                                            # :436-447
        a, b = distill_image(anc[i], sc_old[i], sb[i], cls_inds[i], box_inds[i], tc[i], tb[i],
                             dist_loss_weight, ori, reg_max, report=report)
        d_cls.append(a); d_box.append(b)
    return dict(loss_cls=l_cls, loss_bbox=l_bbox, loss_dfl=l_dfl, loss_dist_cls=d_cls, loss_dist_bbox=d_box)


def total_loss(losses: dict) -> Tensor:
    """mmengine ``parse_losses`` (external): every key containing 'loss' -> tensor.mean()
    or sum of means over a list; total = sum."""
    tot = 0
    for k, v in losses.items():
        if 'loss' in k:
            tot = tot + (v.mean() if isinstance(v, Tensor) else sum(x.mean() for x in v))
    return tot


def erd_step(t_cls, t_box, s_cls, s_box, gt_bboxes, gt_labels, pad_shapes, ori,
             dist_loss_weight=1.0, num_classes=80, reg_max=16, report=None, reduce_mean=None):
    """ERS selection + loss_by_feat + backward: the unit the benchmark calls one step
    (gfl_increment_erd.py:202-220 minus the conv stacks).  Student tensors must be leaf
    tensors with requires_grad.  Returns (losses dict, cls_inds, box_inds)."""
    cls_inds, box_inds = sel_pos(t_cls, t_box, report)
    losses = loss_by_feat(t_cls, t_box, s_cls, s_box, cls_inds, box_inds, ori, dist_loss_weight,
                          gt_bboxes, gt_labels, pad_shapes, num_classes, reg_max, reduce_mean=reduce_mean,
                          report=report)
    total_loss(losses).backward()
    return losses, cls_inds, box_inds
