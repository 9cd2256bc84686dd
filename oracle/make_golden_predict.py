"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/predict_*.pt from the REAL reference
(GFLHead.predict_by_feat executed by path, oracle/ref_by_path.py) for the inference
post-process cases below.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden_predict

Fixtures hold the make_batch arguments (inputs are regenerated from the seed) and the
reference's detections per image."""
from __future__ import annotations

import os

import torch

from erd_b200.synth import make_batch
from oracle import ref_by_path as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# name -> (make_batch kwargs, logit shift added to the student class logits so that a realistic
# number of scores clears score_thr = 0.05; the head's bias init puts them near sigmoid(-4.6))
PREDICT_CASES = {
    'many': (dict(num_imgs=2, img_hw=(512, 640), ori=40, seed=201, mode='trained'), 3.0),
    'odd_size': (dict(num_imgs=2, img_hw=(333, 500), ori=40, seed=202, mode='trained'), 4.0),
    'few': (dict(num_imgs=3, img_hw=(256, 320), ori=40, seed=203), -2.0),
    'none': (dict(num_imgs=1, img_hw=(256, 320), ori=40, seed=204), -6.0),
}


def case_inputs(name):
    kw, shift = PREDICT_CASES[name]
    b = make_batch(**kw)
    return b, [t + shift for t in b.s_cls], b.s_box


def main():
    R.load_reference()
    head = R.build_reference_head(80)
    head.test_cfg = R.ConfigDict(nms_pre=1000, min_bbox_size=0, score_thr=0.05,
                                 nms=dict(type='nms', iou_threshold=0.6), max_per_img=100)
    for name in PREDICT_CASES:
        b, s_cls, s_box = case_inputs(name)
        metas = [dict(img_shape=s, pad_shape=p, scale_factor=(1.0, 1.0)) for s, p in zip(b.img_shapes, b.pad_shapes)]
        out = head.predict_by_feat(s_cls, s_box, batch_img_metas=metas, rescale=False)
        fix = dict(case=name, dets=[dict(bboxes=r.bboxes.clone(), scores=r.scores.clone(), labels=r.labels.clone())
                                    for r in out])
        path = os.path.join(ROOT, 'tests', 'golden', f'predict_{name}.pt')
        torch.save(fix, path)
        print(name, [int(r.bboxes.shape[0]) for r in out], os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
