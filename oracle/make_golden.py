"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.pt from the REAL reference.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden            # all cases
    python -m oracle.make_golden tiny_40_40 # one case

The reference's own ``GFLHeadIncrementERD.loss_by_feat`` and
``GFLIncrementERD.sel_pos`` execute unmodified (``oracle/ref_by_path.py``); the only
restated piece is the third-party ``mmcv.ops.batched_nms``.  Integer outputs are stored
in full, gradients as digests (``oracle/golden_cases.py``) except for the small cases.
"""
from __future__ import annotations

import os
import sys

import torch

from oracle import erd_oracle, ref_by_path
from oracle.golden_cases import CASES, GOLDEN_DIR, case_batch, grad_digest


def run_reference(batch, record=None):
    """One ERS + loss_by_feat + backward through the real reference classes."""
    ref = ref_by_path.load_reference()
    head = ref_by_path.build_reference_head(batch.num_classes, batch.reg_max)
    det = ref_by_path.build_reference_detector_stub(batch.ori, batch.reg_max)
    rec = dict(gt_inds=[], keep=[]) if record is None else record

    orig_assign = head.assigner.assign
    orig_single = head._get_targets_single
    state = {}

    def assign(*a, **k):
        res = orig_assign(*a, **k)
        full = torch.full(state['valid'].shape, -1, dtype=torch.long)
        full[state['valid']] = res.gt_inds
        rec['gt_inds'].append(full)
        return res

    def single(flat_anchors, valid_flags, *a, **k):
        state['valid'] = valid_flags.bool()
        return orig_single(flat_anchors, valid_flags, *a, **k)
    head.assigner.assign = assign
    head._get_targets_single = single

    mod = sys.modules['mmdet.models.dense_heads.gfl_head_increment_erd']
    orig_nms = mod.batched_nms

    def nms_rec(*a, **k):
        out = orig_nms(*a, **k)
        rec['keep'].append(out[1].clone())
        return out
    mod.batched_nms = nms_rec
    try:
        s_cls = [t.clone().requires_grad_() for t in batch.s_cls]
        s_box = [t.clone().requires_grad_() for t in batch.s_box]
        cls_inds, _, box_inds, _ = det.sel_pos(batch.t_cls, batch.t_box)
        gts = [ref.InstanceData(bboxes=b, labels=l) for b, l in zip(batch.gt_bboxes, batch.gt_labels)]
        metas = [dict(img_shape=i, pad_shape=p) for i, p in zip(batch.img_shapes, batch.pad_shapes)]
        losses = head.loss_by_feat((batch.t_cls, batch.t_box), (s_cls, s_box), cls_inds, None, box_inds, None,
                                   batch.ori, 1, None, gts, metas)
        erd_oracle.total_loss(losses).backward()
    finally:
        mod.batched_nms = orig_nms
    return dict(losses={k: [float(x) for x in v] for k, v in losses.items()},
                cls_inds=cls_inds, box_inds=box_inds, keep=rec['keep'], gt_inds=rec['gt_inds'],
                g_cls=[t.grad for t in s_cls], g_box=[t.grad for t in s_box])


def make(name: str):
    kwargs, full = CASES[name]
    batch = case_batch(name)
    out = run_reference(batch)
    gold = dict(name=name, kwargs=kwargs, torch_version=torch.__version__, losses=out['losses'],
                cls_inds=out['cls_inds'], box_inds=out['box_inds'], keep=out['keep'],
                pos=[(g > 0).nonzero().squeeze(1) for g in out['gt_inds']],
                pos_gt=[g[g > 0] for g in out['gt_inds']],
                num_invalid=[int((g < 0).sum()) for g in out['gt_inds']],
                g_cls=[grad_digest(t, 100 + i) for i, t in enumerate(out['g_cls'])],
                g_box=[grad_digest(t, 200 + i) for i, t in enumerate(out['g_box'])])
    if full:
        gold['inputs'] = dict(t_cls=batch.t_cls, t_box=batch.t_box, s_cls=batch.s_cls, s_box=batch.s_box,
                              gt_bboxes=batch.gt_bboxes, gt_labels=batch.gt_labels)
        gold['g_cls_full'] = out['g_cls']
        gold['g_box_full'] = out['g_box']
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + '.pt')
    torch.save(gold, path)
    npos = [int(p.numel()) for p in gold['pos']]
    print(f'{name}: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB) pos={npos} '
          f'K_cls={[len(x) for x in gold["cls_inds"]]} K_box={[len(x) for x in gold["box_inds"]]} '
          f'keep={[len(x) for x in gold["keep"]]}')


if __name__ == '__main__':
    torch.manual_seed(0)
    for nm in (sys.argv[1:] or list(CASES)):
        make(nm)
