"""Shared helpers for the parity tests: run the CUDA path / the oracle on one Batch."""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from erd_b200.synth import Batch  # noqa: E402

LOSS_KEYS = ('loss_cls', 'loss_bbox', 'loss_dfl', 'loss_dist_cls', 'loss_dist_bbox')


def split_losses(vec: torch.Tensor, n_imgs: int, n_levels: int = 5) -> dict:
    v = vec.detach().cpu().tolist()
    L = n_levels
    return dict(loss_cls=v[0:L], loss_bbox=v[L:2 * L], loss_dfl=v[2 * L:3 * L],
                loss_dist_cls=v[3 * L:3 * L + n_imgs], loss_dist_bbox=v[3 * L + n_imgs:3 * L + 2 * n_imgs])


def run_oracle(batch: Batch, dist_loss_weight: float = 1.0) -> dict:
    """CPU oracle (oracle/erd_oracle.py) on the batch: losses, index sets, gradients."""
    from oracle import erd_oracle as O
    s_cls = [t.clone().requires_grad_() for t in batch.s_cls]
    s_box = [t.clone().requires_grad_() for t in batch.s_box]
    rep = {}
    losses, cls_inds, box_inds = O.erd_step(batch.t_cls, batch.t_box, s_cls, s_box, batch.gt_bboxes,
                                            batch.gt_labels, batch.pad_shapes, batch.ori, dist_loss_weight,
                                            batch.num_classes, batch.reg_max, report=rep)
    return dict(losses={k: [float(x) for x in v] for k, v in losses.items()}, cls_inds=cls_inds,
                box_inds=box_inds, keep=rep['keep'], gt_inds=rep['gt_inds'], avg=rep['avg_factors'],
                g_cls=[t.grad for t in s_cls], g_box=[t.grad for t in s_box], report=rep)


def run_cuda(batch: Batch, dist_loss_weight: float = 1.0, path=None, upstream=None) -> dict:
    """The product path through the C ABI on cuda:0; everything returned on the CPU."""
    from erd_b200.ops import ErdPath
    path = path or ErdPath()
    b = batch.to('cuda')
    plan, losses, g_cls, g_box = path.step(b.t_cls, b.t_box, b.s_cls, b.s_box, b.gt_bboxes, b.gt_labels,
                                           b.pad_shapes, b.num_classes, b.ori, b.reg_max, dist_loss_weight,
                                           upstream=upstream)
    torch.cuda.synchronize()
    n = batch.num_imgs
    cc, bc, kc = plan.cls_count.cpu(), plan.box_count.cpu(), plan.keep_count.cpu()
    return dict(
        losses=split_losses(losses, n), loss_vec=losses.cpu(),
        cls_inds=[plan.cls_inds[i, :cc[i]].cpu().long() for i in range(n)],
        box_inds=[plan.box_inds[i, :bc[i]].cpu().long() for i in range(n)],
        keep=[plan.keep[i, :kc[i]].cpu().long() for i in range(n)],
        gt_inds=[plan.gt_inds[i].cpu().long() for i in range(n)],
        num_pos=plan.num_pos.cpu(), avg=tuple(plan.avg.cpu().tolist()), thr=plan.thr.cpu(),
        g_cls=[t.cpu() for t in g_cls], g_box=[t.cpu() for t in g_box], plan=plan)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| relative to the reference tensor's scale."""
    a, b = a.double(), b.double()
    scale = float(b.abs().max())
    return float((a - b).abs().max()) / (scale if scale > 0 else 1.0)


def elem_rel_err(a: torch.Tensor, b: torch.Tensor, floor: float) -> float:
    """max elementwise |a-b| / max(|b|, floor)."""
    a, b = a.double(), b.double()
    return float(((a - b).abs() / b.abs().clamp(min=floor)).max())
