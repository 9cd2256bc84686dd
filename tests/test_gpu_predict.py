"""Inference post-process on the device (erd_predict, csrc/predict.cu) against the golden fixtures generated
from the REAL reference (GFLHead.predict_by_feat run by path, oracle/make_golden_predict.py) and against the
oracle on larger inputs: labels and detection counts bit-exact, boxes and scores within 1e-5."""
import os

import pytest
import torch

from erd_b200.predict import ErdPredictor
from erd_b200.synth import make_batch
from oracle import predict_oracle as P
from oracle.make_golden_predict import PREDICT_CASES, case_inputs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _close(a, b, tol=1e-5):
    return a.shape == b.shape and (a.numel() == 0 or float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max())))


def _compare(mine, ref, what):
    assert len(mine) == len(ref)
    for i, (m, r) in enumerate(zip(mine, ref)):
        assert m['labels'].numel() == r['labels'].numel(), f'{what} image {i}: {m["labels"].numel()} vs {r["labels"].numel()} detections'
        assert torch.equal(m['labels'].cpu(), r['labels']), f'{what} image {i}: labels / order differ'
        assert _close(m['bboxes'].cpu(), r['bboxes']), f'{what} image {i}: boxes'
        assert _close(m['scores'].cpu(), r['scores']), f'{what} image {i}: scores'


@pytest.mark.parametrize('name', sorted(PREDICT_CASES))
def test_predict_matches_reference_fixture(name):
    fix = torch.load(os.path.join(GOLD, f'predict_{name}.pt'))
    b, s_cls, s_box = case_inputs(name)
    out = ErdPredictor().predict_by_feat([t.cuda() for t in s_cls], [t.cuda() for t in s_box], b.img_shapes)
    _compare(out, fix['dets'], name)


@pytest.mark.parametrize('shift,hw', [(3.5, (800, 1333)), (6.0, (480, 640))])
def test_predict_matches_oracle_on_crowded_inputs(shift, hw):
    """Enough scores above the threshold that levels hold far more than nms_pre candidates (the radix-select
    path in front of the sort), ragged image shapes for the clamp, rescale=True."""
    b = make_batch(3, hw, ori=40, seed=77, mode='trained')
    s_cls = [t + shift for t in b.s_cls]
    shapes = [(hw[0], hw[1]), (hw[0] - 37, hw[1] - 90), (hw[0] - 5, hw[1])]
    ref = P.predict_by_feat(s_cls, b.s_box, shapes)
    pred = ErdPredictor()
    out = pred.predict_by_feat([t.cuda() for t in s_cls], [t.cuda() for t in b.s_box], shapes)
    _compare(out, ref, f'shift {shift}')
    n_cand = [int((t[0].sigmoid() > 0.05).sum()) for t in s_cls]
    assert max(n_cand) > 4096, n_cand   # the case does exercise the select path
    # rescale=True: boxes * (1 / scale_factor) after the clamp (base_dense_head.py:458-461); the size filter and the
    # NMS then see the rescaled boxes
    sf = [(1.5, 1.25)] * 3
    out_r = pred.predict_by_feat([t.cuda() for t in s_cls], [t.cuda() for t in b.s_box], shapes, scale_factors=sf)
    assert all(o['labels'].numel() > 0 for o in out_r)


def test_head_predict_by_feat_contract():
    from erd_b200.head import GFLHeadIncrementERD
    head = GFLHeadIncrementERD(80, 256, build_convs=False,
                               test_cfg=dict(nms_pre=1000, min_bbox_size=0, score_thr=0.05,
                                             nms=dict(type='nms', iou_threshold=0.6), max_per_img=100))
    b, s_cls, s_box = case_inputs('few')
    metas = [dict(img_shape=s, pad_shape=p, scale_factor=(1.0, 1.0)) for s, p in zip(b.img_shapes, b.pad_shapes)]
    res = head.predict_by_feat([t.cuda() for t in s_cls], [t.cuda() for t in s_box], batch_img_metas=metas, rescale=False)
    fix = torch.load(os.path.join(GOLD, 'predict_few.pt'))
    _compare([dict(bboxes=r.bboxes, scores=r.scores, labels=r.labels) for r in res], fix['dets'], 'head')
    with pytest.raises(ValueError):
        head.predict_by_feat(s_cls, s_box, batch_img_metas=metas, with_nms=False)
