"""Parity soak: index parity of ERS / ATSS / teacher NMS over several hundred images (both class
splits, iid and planted-object teachers, ragged pads), with every mismatch CLASSIFIED from the
oracle's margins instead of skipped (the CUDA scan uses exp-based
arithmetic whose last bit may differ from torch-CPU's, which can only flip a decision that sits
within a few ulp of its threshold).

A mismatch is a *tie-class* difference when the oracle's own margin at that decision is below
TIE_ULPS ulp; anything else fails the test.  The summary line is printed either way, so a run with
zero mismatches documents that as well."""
import math

import pytest
import torch

from erd_b200.ops import ErdPath
from erd_b200.synth import make_batch
from oracle import erd_oracle as O
from util import run_cuda

pytestmark = pytest.mark.gpu

TIE_ULPS = 4.0
IMGS_PER_BATCH = 16
BATCHES = 32          # 512 images


def _ulp(x: float) -> float:
    return abs(float(torch.nextafter(torch.tensor(x, dtype=torch.float32), torch.tensor(math.inf)) - x))


def _oracle_indices(batch):
    """Forward-only oracle: ERS lists, assigned_gt_inds, NMS keep lists, and the margins behind them."""
    rep = {}
    with torch.no_grad():
        cls_inds, box_inds = O.sel_pos(batch.t_cls, batch.t_box, rep)
        O.loss_by_feat(batch.t_cls, batch.t_box, batch.s_cls, batch.s_box, cls_inds, box_inds, batch.ori, 1.0,
                       batch.gt_bboxes, batch.gt_labels, batch.pad_shapes, batch.num_classes, batch.reg_max,
                       report=rep)
    return cls_inds, box_inds, rep


def test_index_parity_soak_512_images():
    path = ErdPath()
    counts = dict(images=0, ers=0, atss=0, nms=0, ers_tie=0, atss_tie=0, nms_after_ers=0)
    hard = []
    min_margin_ulps = dict(cls=math.inf, box=math.inf)
    for k in range(BATCHES):
        mode = 'trained' if k % 2 else 'gaussian'
        ori = 70 if k % 4 >= 2 else 40
        hw = [(800, 1333), (768, 1024), (640, 960), (1024, 1024)][(k // 4) % 4]
        ch, cw = (math.ceil(hw[0] / 32) * 32, math.ceil(hw[1] / 32) * 32)
        pads = None
        if k % 3 == 0:   # ragged batch: some images smaller than the canvas
            pads = [(ch - 32 * ((i * 7 + k) % 4), cw - 32 * ((i * 5 + k) % 5)) for i in range(IMGS_PER_BATCH)]
        batch = make_batch(IMGS_PER_BATCH, hw, ori=ori, seed=7000 + k, mode=mode, gt_size_pow=2.0,
                           num_gt=(0, 12) if k % 5 == 0 else None, pad_shapes=pads)
        c = run_cuda(batch, path=path)
        o_cls, o_box, rep = _oracle_indices(batch)
        for i in range(batch.num_imgs):
            counts['images'] += 1
            thr_c, thr_b = rep['cls_thr'][i], rep['box_thr'][i]
            mc, mb = rep['cls_margin'][i] / _ulp(thr_c), rep['box_margin'][i] / _ulp(thr_b)
            min_margin_ulps['cls'] = min(min_margin_ulps['cls'], mc)
            min_margin_ulps['box'] = min(min_margin_ulps['box'], mb)
            ers_ok = torch.equal(c['cls_inds'][i], o_cls[i]) and torch.equal(c['box_inds'][i], o_box[i])
            if not ers_ok:
                counts['ers'] += 1
                if min(mc, mb) <= TIE_ULPS:
                    counts['ers_tie'] += 1
                else:
                    hard.append(f'batch {k} image {i}: ERS lists differ, oracle margins {mc:.1f} / {mb:.1f} ulp')
            if not torch.equal(c['gt_inds'][i], rep['gt_inds'][i]):
                counts['atss'] += 1
                a = rep['atss'][i]
                thr_margin = a.get('min_thr_margin', math.inf)
                if a.get('topk_boundary_ties') or thr_margin <= TIE_ULPS * 1.2e-7:
                    counts['atss_tie'] += 1
                else:
                    hard.append(f'batch {k} image {i}: assigned_gt_inds differ, IoU-threshold margin {thr_margin:.3g}, '
                                f'no top-k distance tie')
            if not torch.equal(c['keep'][i], rep['keep'][i]):
                counts['nms'] += 1
                if not ers_ok:
                    counts['nms_after_ers'] += 1   # different candidate set: follows from the ERS difference
                else:
                    hard.append(f'batch {k} image {i}: NMS keep lists differ on identical candidates '
                                f'({len(c["keep"][i])} vs {len(rep["keep"][i])} kept, '
                                f'{rep["score_ties"][i]} score ties in the oracle)')
    print(f'\nsoak: {counts}  smallest oracle margins: cls {min_margin_ulps["cls"]:.1f} ulp, '
          f'box {min_margin_ulps["box"]:.1f} ulp')
    assert counts['images'] >= 500
    assert not hard, '\n'.join(hard)


def test_losses_with_world_averaged_factors():
    """Two ranks' worth of different batches on one device: each 'rank' runs the assignment chain, the two
    avg factors are averaged as reduce_mean would (t / W summed in rank order, dist_utils.py:59-65,
    gfl_head_increment_erd.py:390-391,406-407), and each rank's losses and gradients are compared with the
    oracle fed the same world-averaged factors through its reduce_mean hook."""
    path = ErdPath()
    batches = [make_batch(2, (512, 640), ori=40, seed=910 + r, mode='trained' if r else 'gaussian', gt_size_pow=2.0,
                          num_gt=[1, 7] if r else [9, 3]) for r in range(2)]
    dev = [b.to('cuda') for b in batches]
    plans, outs = [], []
    for b in dev:   # the two "ranks" need separate plans: same geometry, so force distinct paths
        p = ErdPath().plan(b.s_cls, b.num_classes, b.ori, b.reg_max)
        p.set_targets(b.gt_bboxes, b.gt_labels, b.pad_shapes)
        path.prepare(p, b.t_cls, b.t_box, b.s_cls, b.s_box)
        plans.append(p)
    torch.cuda.synchronize()
    local = [p.avg.clone() for p in plans]
    mean = local[0] / 2 + local[1] / 2          # div_(world) then SUM in rank order
    for p, b in zip(plans, dev):
        p.avg.copy_(mean)
        g_cls = [torch.empty_like(t) for t in b.s_cls]
        g_box = [torch.empty_like(t) for t in b.s_box]
        losses = torch.empty(p.num_losses, device='cuda')
        path.loss_fwd_bwd(p, b.t_cls, b.t_box, b.s_cls, b.s_box, g_cls, g_box, losses, 1.0)
        torch.cuda.synchronize()
        outs.append((losses.cpu(), [t.cpu() for t in g_cls], [t.cpu() for t in g_box]))
    # oracle: first pass for the local factors, second pass with the mean over "ranks"
    loc = []
    for b in batches:
        rep = {}
        with torch.no_grad():
            ci, bi = O.sel_pos(b.t_cls, b.t_box)
            O.loss_by_feat(b.t_cls, b.t_box, b.s_cls, b.s_box, ci, bi, b.ori, 1.0, b.gt_bboxes, b.gt_labels,
                           b.pad_shapes, b.num_classes, b.reg_max, report=rep)
        loc.append(rep['avg_local'])
    for r, b in enumerate(batches):
        assert abs(float(local[r][0]) - loc[r][0]) < 1e-6 and abs(float(local[r][1]) - loc[r][1]) <= 1e-5 * max(1.0, loc[r][1])
        calls = []

        def reduce_mean(t, calls=calls):
            idx = len(calls)
            calls.append(float(t))
            return torch.tensor(loc[0][idx] / 2, dtype=torch.float) + torch.tensor(loc[1][idx] / 2, dtype=torch.float)
        s_cls = [t.clone().requires_grad_() for t in b.s_cls]
        s_box = [t.clone().requires_grad_() for t in b.s_box]
        losses, _, _ = O.erd_step(b.t_cls, b.t_box, s_cls, s_box, b.gt_bboxes, b.gt_labels, b.pad_shapes, b.ori,
                                  1.0, b.num_classes, b.reg_max, reduce_mean=reduce_mean)
        assert len(calls) == 2
        flat = losses['loss_cls'] + losses['loss_bbox'] + losses['loss_dfl'] + losses['loss_dist_cls'] + losses['loss_dist_bbox']
        got, g_cls, g_box = outs[r]
        for j, x in enumerate(flat):
            assert abs(float(got[j]) - float(x)) <= 1e-5 * max(abs(float(x)), 1e-7), (r, j, float(got[j]), float(x))
        for a, ref in zip(g_cls + g_box, [t.grad for t in s_cls + s_box]):
            scale = float(ref.abs().max())
            assert float((a - ref).abs().max()) <= 1e-5 * max(scale, 1e-12)
