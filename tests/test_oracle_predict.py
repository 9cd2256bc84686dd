"""Oracle of the NEXT scope row (inference post-process) pinned to the
reference: against golden fixtures generated from the real GFLHead.predict_by_feat, and --
where the reference tree exists -- against the reference itself.  CPU only."""
import os

import pytest
import torch

from oracle import predict_oracle as P
from oracle import ref_by_path as R
from oracle.make_golden_predict import PREDICT_CASES, case_inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('name', sorted(PREDICT_CASES))
def test_oracle_matches_reference_fixture(name):
    fix = torch.load(os.path.join(GOLD, f'predict_{name}.pt'))
    b, s_cls, s_box = case_inputs(name)
    out = P.predict_by_feat(s_cls, s_box, b.img_shapes)
    assert len(out) == len(fix['dets'])
    for mine, ref in zip(out, fix['dets']):
        assert torch.equal(mine['labels'], ref['labels'])
        assert torch.equal(mine['bboxes'], ref['bboxes'])
        assert torch.equal(mine['scores'], ref['scores'])


def test_fixture_shapes_cover_the_edge_cases():
    n = {name: [int(d['bboxes'].shape[0]) for d in torch.load(os.path.join(GOLD, f'predict_{name}.pt'))['dets']]
         for name in PREDICT_CASES}
    assert all(k == 100 for k in n['many'])            # max_per_img truncation
    assert any(0 < k < 100 for k in n['few'])          # fewer survivors than max_per_img
    assert n['none'] == [0]                            # nothing clears score_thr


@pytest.mark.skipif(not R.available(), reason='reference tree not present')
def test_oracle_matches_reference():
    R.load_reference()
    head = R.build_reference_head(80)
    head.test_cfg = R.ConfigDict(nms_pre=1000, min_bbox_size=0, score_thr=0.05,
                                 nms=dict(type='nms', iou_threshold=0.6), max_per_img=100)
    from erd_b200.synth import make_batch
    for seed, hw, shift in [(31, (480, 640), 3.0), (32, (800, 1333), 2.5)]:
        b = make_batch(1, hw, ori=40, seed=seed, mode='trained')
        s_cls = [t + shift for t in b.s_cls]
        metas = [dict(img_shape=s, pad_shape=p, scale_factor=(1.0, 1.0)) for s, p in zip(b.img_shapes, b.pad_shapes)]
        ref = head.predict_by_feat(s_cls, b.s_box, batch_img_metas=metas, rescale=False)
        mine = P.predict_by_feat(s_cls, b.s_box, b.img_shapes)
        for r, m in zip(ref, mine):
            assert torch.equal(r.bboxes, m['bboxes']) and torch.equal(r.scores, m['scores'])
            assert torch.equal(r.labels, m['labels'])
